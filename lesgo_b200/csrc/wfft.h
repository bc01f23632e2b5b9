// wfft.h -- warp-scope Stockham FFT.
//
// One warp transforms NF rows of length M (complex) in its PRIVATE slice of shared memory, so
// the only synchronisation on the path is __syncwarp(): warps of a block never wait for each
// other and the scheduler always has independent warps to hide shared/global latency with
// (the block-cooperative passes of round 1 spent a third of their stall cycles in barriers).
// The stage arithmetic (radices, twiddle layout, index maps) is that of fft_core.h.
//
//   * M*NF <= 256: every lane holds <= 8 points, stages run IN PLACE on one padded buffer
//     (all loads, __syncwarp, all stores) and the stage twiddles, which depend only on the
//     lane, live in REGISTERS for the whole persistent loop (no twiddle LDS at all).
//   * larger rows ping-pong between two buffers (one butterfly at a time: load, twiddle,
//     DFT, store), twiddles read from the per-stage shared-memory tables.
#pragma once
#include <type_traits>
#include "fft_core.h"

namespace lg {

#ifdef LESGO_EMUL
#define LG_SYNCWARP() emu::syncthreads()   // fibres advance one sync point per round: warp-exact
#else
#define LG_SYNCWARP() __syncwarp()
#endif

// lane-resident twiddles of one stage (radix R, Ns = product of earlier radices)
template <int M, int NF, int R, int Ns>
struct WTw {
    static constexpr int T = M / R, ITEMS = NF * T, IPT = (ITEMS + 31) / 32;
    cplx w[IPT][R > 1 ? R - 1 : 1];
    LG_D void init(const cplx* __restrict__ Wst, int lane) {
        if constexpr (Ns > 1 && R > 1) {
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const int j = (lane + 32 * q) % T, k = j % Ns;
#pragma unroll
                for (int r = 1; r < R; ++r) w[q][r - 1] = Wst[(r - 1) * Ns + k];
            }
        }
    }
};

// A store functor may offer pre(f, i, slot) / put(f, i, slot, v) next to operator(): the stage then calls pre for
// the R outputs of a butterfly BEFORE it computes them and put afterwards, so an epilogue that reads operands
// (EpiFused) has all its loads in flight at once instead of one load-use-store chain per element.
template <class T, class = void> struct HasPre : std::false_type {};
template <class T> struct HasPre<T, std::void_t<decltype(&T::pre)>> : std::true_type {};

// One stage over the warp's NF rows.
//   SRC / DST = 0: functor ld(f, i) / st(f, i, v);  1: padded shared buffer (row f at f*SL)
//   MIDSYNC   : source and destination are the same buffer -> hold every butterfly of the lane
//               in registers across a __syncwarp
//   TWREG     : twiddles from tw (registers) instead of the shared table Wst
template <int M, int NF, int R, int Ns, bool INV, int SRC, int DST, bool MIDSYNC, bool TWREG, class Ld, class St>
LG_D void wstage(int lane, const cplx* sbuf, cplx* dbuf, const cplx* __restrict__ Wst,
                 const WTw<M, NF, R, Ns>& tw, Ld ld, St st) {
    constexpr int T = M / R, ITEMS = NF * T, IPT = (ITEMS + 31) / 32, SL = SmemLen<M>::value;
    constexpr int NV = MIDSYNC ? IPT : 1;
    cplx v[NV][R];
    auto store = [&](int f, int j, const cplx* vv) {
        const int j0 = (Ns == 1) ? j * R : ((j / Ns) * Ns * R + (j % Ns));
        if constexpr (DST == 1) {
            if constexpr (Ns % 8 == 0 || (Ns == 1 && (R == 8 || R == 4 || R == 2))) {
                cplx* p = dbuf + f * SL + spad(j0);
                constexpr int step = (Ns % 8 == 0) ? (Ns + Ns / 8) : 1;
#pragma unroll
                for (int r = 0; r < R; ++r) p[r * step] = vv[r];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) dbuf[f * SL + spad(j0 + r * Ns)] = vv[r];
            }
        } else if constexpr (HasPre<St>::value) {
#pragma unroll
            for (int r = 0; r < R; ++r) st.put(f, j0 + r * Ns, r, vv[r]);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) st(f, j0 + r * Ns, vv[r]);
        }
    };
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int it = lane + 32 * q;
        cplx* vv = v[MIDSYNC ? q : 0];
        if (ITEMS % 32 == 0 || it < ITEMS) {
            const int f = (NF == 1) ? 0 : it / T, j = (NF == 1) ? it : it % T;
            if constexpr (DST == 0 && HasPre<St>::value) {
                const int j0 = (Ns == 1) ? j * R : ((j / Ns) * Ns * R + (j % Ns));
#pragma unroll
                for (int r = 0; r < R; ++r) st.pre(f, j0 + r * Ns, r);
            }
            if constexpr (SRC == 1) {
                if constexpr (T % 8 == 0) {
                    const cplx* p = sbuf + f * SL + spad(j);
#pragma unroll
                    for (int r = 0; r < R; ++r) vv[r] = p[r * (T + T / 8)];
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) vv[r] = sbuf[f * SL + spad(j + r * T)];
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) vv[r] = ld(f, j + r * T);
            }
            if constexpr (Ns > 1) {
                const int k = j % Ns;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const cplx w = TWREG ? tw.w[q][r - 1] : Wst[(r - 1) * Ns + k];
                    vv[r] = INV ? cmulc(vv[r], w) : cmul(vv[r], w);
                }
            }
            if (INV) {
#pragma unroll
                for (int r = 0; r < R; ++r) vv[r] = cswap(vv[r]);
            }
            Dft<R>::run(vv);
            if (INV) {
#pragma unroll
                for (int r = 0; r < R; ++r) vv[r] = cswap(vv[r]);
            }
            if constexpr (!MIDSYNC) store(f, j, vv);
        }
    }
    if constexpr (MIDSYNC) {
        LG_SYNCWARP();
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int it = lane + 32 * q;
            if (ITEMS % 32 == 0 || it < ITEMS) {
                const int f = (NF == 1) ? 0 : it / T, j = (NF == 1) ? it : it % T;
                store(f, j, v[q]);
            }
        }
    }
}

// The whole transform.  A, B: the warp's buffers (B == A when INPLACE).
//   FROM_SMEM : the input already sits in A (natural order, padded);  else ld(f, i)
//   TO_SMEM   : the natural-order result is left in a buffer (returned; caller syncs before
//               reading it);  else it is delivered through st(f, i, v)
template <int M, int NF, bool INV, bool INPLACE, bool TWREG>
struct WarpFft {
    typedef Plan<M> P;
    typedef PlanInfo<M> PI;
    static constexpr int R1 = P::R1, R2 = P::R2, R3 = P::R3, R4 = P::R4, NST = PI::nstages;
    WTw<M, NF, R2, R1> t2;
    WTw<M, NF, R3, R1 * R2> t3;
    WTw<M, NF, R4, R1 * R2 * R3> t4;
    WTw<M, NF, R1, 1> t1;   // empty (first stage has no twiddles)

    LG_D void init(const cplx* __restrict__ W, int lane) {
        if constexpr (TWREG) {
            if constexpr (NST >= 2) t2.init(W + PI::off2, lane);
            if constexpr (NST >= 3) t3.init(W + PI::off3, lane);
            if constexpr (NST >= 4) t4.init(W + PI::off4, lane);
        }
    }

    template <bool FROM_SMEM, bool TO_SMEM, class Ld, class St>
    LG_D cplx* run(cplx* A, cplx* B, const cplx* __restrict__ W, int lane, Ld ld, St st) const {
        constexpr int S0 = FROM_SMEM ? 1 : 0;
        // buffer written by stage s (1-based) when it writes shared memory
        cplx* d1 = FROM_SMEM ? B : A;                 // INPLACE: B == A
        cplx* d2 = (d1 == A) ? B : A;
        cplx* d3 = (d2 == A) ? B : A;
        cplx* d4 = (d3 == A) ? B : A;
        constexpr bool MS = INPLACE;                  // smem -> same smem
        if constexpr (NST == 1) {
            wstage<M, NF, R1, 1, INV, S0, TO_SMEM ? 1 : 0, FROM_SMEM && MS, TWREG>(lane, A, d1, W, t1, ld, st);
            return TO_SMEM ? d1 : nullptr;
        } else {
            wstage<M, NF, R1, 1, INV, S0, 1, FROM_SMEM && MS, TWREG>(lane, A, d1, W, t1, ld, st);
            LG_SYNCWARP();
            if constexpr (NST == 2) {
                wstage<M, NF, R2, R1, INV, 1, TO_SMEM ? 1 : 0, TO_SMEM && MS, TWREG>(lane, d1, d2, W + PI::off2, t2, ld, st);
                return TO_SMEM ? d2 : nullptr;
            } else {
                wstage<M, NF, R2, R1, INV, 1, 1, MS, TWREG>(lane, d1, d2, W + PI::off2, t2, ld, st);
                LG_SYNCWARP();
                if constexpr (NST == 3) {
                    wstage<M, NF, R3, R1 * R2, INV, 1, TO_SMEM ? 1 : 0, TO_SMEM && MS, TWREG>(lane, d2, d3, W + PI::off3, t3, ld, st);
                    return TO_SMEM ? d3 : nullptr;
                } else {
                    wstage<M, NF, R3, R1 * R2, INV, 1, 1, MS, TWREG>(lane, d2, d3, W + PI::off3, t3, ld, st);
                    LG_SYNCWARP();
                    wstage<M, NF, R4, R1 * R2 * R3, INV, 1, TO_SMEM ? 1 : 0, TO_SMEM && MS, TWREG>(lane, d3, d4, W + PI::off4, t4, ld, st);
                    return TO_SMEM ? d4 : nullptr;
                }
            }
        }
    }
};

}  // namespace lg
