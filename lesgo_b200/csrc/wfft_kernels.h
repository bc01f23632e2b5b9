// wfft_kernels.h -- x passes of the batched 2-D real FFT at WARP scope (see wfft.h).
//
// Same contract as k_xfwd / k_xinv of fft_kernels.h (same prologue / epilogue functors, same
// XfOut / XiSrc descriptors), different execution model: each warp of a persistent block owns
// whole rows -- prologue, half-length complex FFT, real<->half-complex untangling, epilogue --
// and synchronises only with itself.
#pragma once
#include "fft_kernels.h"
#include "wfft.h"

// LG_XW_TMA = 1: a staged spectral row arrives by ONE bulk asynchronous copy (cp.async.bulk, completion counted on the
// warp's mbarrier) instead of 16-byte cp.async pieces issued by every lane
#ifndef LG_XW_TMA
#define LG_XW_TMA 1
#endif
#ifdef LESGO_EMUL
#undef LG_XW_TMA
#define LG_XW_TMA 0
#endif


namespace lg {

#ifndef LG_XW_TWREG
#define LG_XW_TWREG 1
#endif

// PREF_: the x-inverse kernel stages the NEXT row of each warp with cp.async while the current one is
// being transformed (one staging row of M+1 spectral columns per warp)
template <int NX, bool TWREG_ = (LG_XW_TWREG != 0), bool PREF_ = false> struct XWCfg {
    static constexpr int M = NX / 2;
    typedef PlanInfo<M> PI;
    static constexpr int TMIN = M / PI::rmax;
    static constexpr int NF = TMIN >= 32 ? 1 : (32 / TMIN > 8 ? 8 : 32 / TMIN);   // rows per warp
    static constexpr bool SMALL = (M * NF <= 256);
    // in place up to 8 points per lane (12-point rows in place -- two butterflies per lane held across the
    // __syncwarp, 128 registers -- measured slower than ping-pong buffers: 2.96 against 2.76 ms)
    static constexpr bool INPLACE = SMALL;
    static constexpr bool TWREG = SMALL && TWREG_;
    static constexpr int SL = SmemLen<M>::value;
    static constexpr bool PREF = PREF_ && NF == 1;
    static constexpr int STG = PREF ? (M + 1) : 0;
    static constexpr int WBUF = (INPLACE ? 1 : 2) * NF * SL + STG;      // cplx per warp
    static constexpr int TWL = PI::twlen, NWH = M / 2 + 1;
    static constexpr size_t smem_for(int wpb) {
        return size_t(wpb * WBUF + TWL + NWH) * sizeof(cplx) + size_t(wpb) * 2 * NF * sizeof(int);
    }
    // warps per block: the count (4..8) that lets the most warps be resident per SM
    static constexpr int warps_per_sm(int wpb) { return wpb * int((227 * 1024) / (smem_for(wpb) + 1024)); }
    static constexpr int best_wpb() {
        int best = 8;
        for (int w = 7; w >= 4; --w)
            if (warps_per_sm(w) > warps_per_sm(best)) best = w;
        return best;
    }
    static constexpr int WPB = smem_for(8) > 226 * 1024 ? 4 : best_wpb();
    static constexpr int NTHR = 32 * WPB;
    static constexpr size_t smem = smem_for(WPB);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int by_regs = TWREG ? 2 : (PREF ? 3 : (M <= 256 ? 4 : 3));   // 128 / 80 / 64 / 80 registers per thread
    static constexpr int MINB = by_regs < by_smem ? by_regs : by_smem;
};

// rows of a warp's tile: (plane, y) of row f, plane < 0 = past the end
template <int NF> struct WRows {
    int* sk; int* sy;      // per-warp shared arrays (NF > 1)
    int k0_, y0_;          // NF == 1: registers
    LG_D void set(int lane, unsigned row0, unsigned nrows, int ny, int k0) {
        if constexpr (NF == 1) {
            k0_ = row0 < nrows ? k0 + int(row0 / unsigned(ny)) : -1;
            y0_ = int(row0 % unsigned(ny));
        } else {
            if (lane < NF) {
                const unsigned r = row0 + lane;
                sk[lane] = r < nrows ? k0 + int(r / unsigned(ny)) : -1;
                sy[lane] = int(r % unsigned(ny));
            }
            LG_SYNCWARP();
        }
    }
    LG_D int k(int f) const { if constexpr (NF == 1) return k0_; else return sk[f]; }
    LG_D int y(int f) const { if constexpr (NF == 1) return y0_; else return sy[f]; }
};

// ---------------------------------------------------------------------------------
// x forward: real rows -> half spectrum rows
// ---------------------------------------------------------------------------------
template <int NX, class Pro, bool TWR = (LG_XW_TWREG != 0)>
__global__ void __launch_bounds__(XWCfg<NX, TWR>::NTHR, XWCfg<NX, TWR>::MINB)
k_xfwd_w(const __grid_constant__ Pro pro, const __grid_constant__ XfOut out, int nfields, int ny, int k0, int nplanes,
         const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef XWCfg<NX, TWR> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    cplx* A = sm + wib * C::WBUF;
    cplx* B = C::INPLACE ? A : A + NF * SL;
    cplx* W = sm + C::WPB * C::WBUF;
    cplx* Wh = W + C::TWL;
    WRows<NF> rows;
    rows.sk = reinterpret_cast<int*>(Wh + C::NWH) + wib * 2 * NF;
    rows.sy = rows.sk + NF;
    load_table(W, Wg, C::TWL);
    load_table(Wh, Whg, C::NWH);
    __syncthreads();
    WarpFft<M, NF, false, C::INPLACE, C::TWREG> fft;
    fft.init(W, lane);

    const unsigned nrows = unsigned(ny) * unsigned(nplanes);
    const unsigned ntiles = (nrows + NF - 1) / NF;
    const unsigned nwork = ntiles * unsigned(nfields);
    const unsigned wstride = gridDim.x * C::WPB;
    // work items dealt round-robin, field index fastest (fields that share inputs run together)
    for (unsigned work = blockIdx.x * C::WPB + wib; work < nwork; work += wstride) {
        const int fld = int(work % unsigned(nfields));
        rows.set(lane, (work / unsigned(nfields)) * NF, nrows, ny, k0);
        cplx* X = fft.template run<false, true>(A, B, W, lane,
            [&](int f, int i) {
                const int k = rows.k(f);
                if (k < 0) return make_double2(0.0, 0.0);
                return pro.load(fld, k, rows.y(f), i);
            },
            [](int, int, cplx) {});
        LG_SYNCWARP();
        // untangle: X_m = E_m + W_N^m O_m,  X_{M-m} = conj(E_m - W_N^m O_m);  m = 0 and M/2 are special
        constexpr int NPM = M / 2 + 1;
        constexpr int ITER = (NF * NPM + 31) / 32;
#pragma unroll
        for (int q = 0; q < ITER; ++q) {
            const int it = lane + 32 * q;
            if (it < NF * NPM) {
                const int f = (NF == 1) ? 0 : it / NPM, m = (NF == 1) ? it : it % NPM;
                const int k = rows.k(f);
                if (k >= 0) {
                    double* drow = out.dst[fld] + poff(k, out.plane, out.ring) + long(rows.y(f)) * out.row;
                    const cplx a = X[f * SL + spad(m)];
                    if (m == 0) {
                        if (out.ncol > 0) *reinterpret_cast<cplx*>(drow) = make_double2(a.x + a.y, 0.0);
                        if (out.write_nyq && out.ncol >= M)
                            *reinterpret_cast<cplx*>(drow + 2 * M) = make_double2(out.write_nyq == 2 ? a.x - a.y : 0.0, 0.0);
                        else if (out.write_nyq && out.ncol < M)
                            *reinterpret_cast<cplx*>(drow + 2 * out.ncol) = make_double2(0.0, 0.0);
                    } else if (m == M / 2) {
                        if (M / 2 < out.ncol) *reinterpret_cast<cplx*>(drow + M) = make_double2(a.x, -a.y);
                    } else {
                        const cplx bz = X[f * SL + spad(M - m)];
                        const cplx b = make_double2(bz.x, -bz.y);
                        const cplx e = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
                        const cplx d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
                        const cplx o = make_double2(d.y, -d.x);               // d / i
                        const cplx t = cmul(o, Wh[m]);
                        if (m < out.ncol) *reinterpret_cast<cplx*>(drow + 2 * m) = cadd(e, t);
                        if (M - m < out.ncol)
                            *reinterpret_cast<cplx*>(drow + 2 * (M - m)) = make_double2(e.x - t.x, -(e.y - t.y));
                    }
                }
            }
        }
        LG_SYNCWARP();
    }
}

// ---------------------------------------------------------------------------------
// x inverse: half spectrum rows -> real rows
// ---------------------------------------------------------------------------------
// Store functor for epilogues with operands (Epi::Ops, Epi::load, Epi::apply): see HasPre in wfft.h
#ifndef LG_EPI_BATCH
#define LG_EPI_BATCH 1
#endif
template <class Epi, class = void> struct EpiBatched : std::false_type {};
template <class Epi> struct EpiBatched<Epi, std::void_t<typename Epi::Ops>> : std::integral_constant<bool, LG_EPI_BATCH != 0> {};
template <class Epi, class Rows>
struct XinvBatchSt {
    const Epi& epi;
    const Rows& rows;
    int fld;
    typename Epi::Ops ops[8];
    LG_D void pre(int f, int i, int s) {
        const int k = rows.k(f);
        if (k >= 0) ops[s] = epi.load(fld, k, rows.y(f), i);
    }
    LG_D void put(int f, int i, int s, cplx v) const {
        const int k = rows.k(f);
        if (k >= 0) epi.apply(fld, k, rows.y(f), i, v, ops[s]);
    }
    LG_D void operator()(int f, int i, cplx v) const {
        const int k = rows.k(f);
        if (k >= 0) epi.store(fld, k, rows.y(f), i, v);
    }
};

// resident blocks asked of the compiler for the epilogues with operands: 2 (128 registers) so that the operands of
// the four outputs of a butterfly stay in registers.  Measured on B200 at 512 x 512 x 256, the four fused x-inverse
// launches of a core step: 5.26 ms unbatched, 6.04 ms batched at 80 registers (spills), 4.90 ms batched at 128.
#ifndef LG_XWF_MINB
#define LG_XWF_MINB 2
#endif
template <class Epi, class C> struct XWMinB {
    static constexpr int value = (EpiBatched<Epi>::value && LG_XWF_MINB > 0 && LG_XWF_MINB < C::MINB) ? LG_XWF_MINB : C::MINB;
};

template <int NX, class Epi, bool TWR = (LG_XW_TWREG != 0), bool PRF = false>
__global__ void __launch_bounds__(XWCfg<NX, TWR, PRF>::NTHR, XWMinB<Epi, XWCfg<NX, TWR, PRF>>::value)
k_xinv_w(const __grid_constant__ XiSrc in, const __grid_constant__ Epi epi, int nfields, int ny, int k0, int nplanes,
         const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef XWCfg<NX, TWR, PRF> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    cplx* A = sm + wib * C::WBUF;
    cplx* B = C::INPLACE ? A : A + NF * SL;
    cplx* W = sm + C::WPB * C::WBUF;
    cplx* Wh = W + C::TWL;
    WRows<NF> rows;
    rows.sk = reinterpret_cast<int*>(Wh + C::NWH) + wib * 2 * NF;
    rows.sy = rows.sk + NF;
    load_table(W, Wg, C::TWL);
    load_table(Wh, Whg, C::NWH);
    __syncthreads();
    WarpFft<M, NF, true, C::INPLACE, C::TWREG> fft;
    fft.init(W, lane);
#if LG_XW_TMA
    __shared__ __align__(8) unsigned long long mbar[C::WPB];     // PREF: one per warp (wfft2_kernels.h: LG_XW_TMA)
    if (C::PREF && lane == 0) mbar_init(&mbar[wib], 1);
    unsigned phase = 0;
    __syncthreads();
#endif

    const unsigned nrows = unsigned(ny) * unsigned(nplanes);
    const unsigned ntiles = (nrows + NF - 1) / NF;
    const unsigned nwork = ntiles * unsigned(nfields);
    const unsigned wstride = gridDim.x * C::WPB;
    cplx* ST = A + (C::INPLACE ? 1 : 2) * NF * SL;           // PREF: this warp's staged spectral row
    auto prefetch = [&](unsigned work) {
        if constexpr (C::PREF) {
#if LG_XW_TMA
            if (lane == 0 && work < nwork) {                 // one bulk copy of the row, counted on the warp's mbarrier
                const int pf = int(work % unsigned(nfields));
                const unsigned r = work / unsigned(nfields);
                const double* srow = in.src[pf] + poff(k0 + int(r / unsigned(ny)), in.plane, in.ring) + long(r % unsigned(ny)) * in.row;
                const int nc = in.ncol < M + 1 ? in.ncol : M + 1;
                mbar_expect_tx(&mbar[wib], unsigned(nc > 0 ? nc : 0) * unsigned(sizeof(cplx)));
                if (nc > 0) bulk_g2s(ST, srow, unsigned(nc) * unsigned(sizeof(cplx)), &mbar[wib]);
            }
#else
            if (work < nwork) {
                const int pf = int(work % unsigned(nfields));
                const unsigned r = work / unsigned(nfields);
                const double* srow = in.src[pf] + poff(k0 + int(r / unsigned(ny)), in.plane, in.ring) + long(r % unsigned(ny)) * in.row;
                const int nc = in.ncol < M + 1 ? in.ncol : M + 1;
                for (int m = lane; m < nc; m += 32) cp_async16(ST + m, srow + 2 * m);
            }
            cp_async_commit();
#endif
        }
    };
    prefetch(blockIdx.x * C::WPB + wib);
    for (unsigned work = blockIdx.x * C::WPB + wib; work < nwork; work += wstride) {
        const int fld = int(work % unsigned(nfields));
        rows.set(lane, (work / unsigned(nfields)) * NF, nrows, ny, k0);
#if LG_XW_TMA
        if constexpr (C::PREF) { mbar_wait(&mbar[wib], phase); phase ^= 1u; LG_SYNCWARP(); }
#else
        if constexpr (C::PREF) { cp_async_wait_all(); LG_SYNCWARP(); }
#endif
        // tangle: Z'_m = E'_m + i O'_m,  E' = X_m + conj(X_{M-m}),  O' = (X_m - conj(X_{M-m})) conj(W_N^m)
        constexpr int NPM = M / 2 + 1;
        constexpr int ITER = (NF * NPM + 31) / 32;
        constexpr int UN = ITER < 4 ? ITER : 4;              // row pairs in flight per lane
#pragma unroll
        for (int q0 = 0; q0 < ITER; q0 += UN) {
            cplx va[UN], vb[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = lane + 32 * (q0 + u);
                va[u] = make_double2(0.0, 0.0); vb[u] = va[u];
                if (q0 + u < ITER && it < NF * NPM) {
                    const int f = (NF == 1) ? 0 : it / NPM, m = (NF == 1) ? it : it % NPM;
                    const int k = rows.k(f);
                    if (C::PREF) {
                        if (m == 0) {
                            va[u].x = in.ncol > 0 ? ST[0].x : 0.0;
                            va[u].y = in.ncol > M ? ST[M].x : 0.0;
                        } else {
                            if (m < in.ncol) va[u] = ST[m];
                            if (m != M / 2 && M - m < in.ncol) vb[u] = ST[M - m];
                        }
                    } else if (k >= 0) {
                        const double* srow = in.src[fld] + poff(k, in.plane, in.ring) + long(rows.y(f)) * in.row;
                        if (m == 0) {                          // real parts of X_0 and X_M only
                            va[u].x = in.ncol > 0 ? ld_cg(srow).x : 0.0;
                            va[u].y = in.ncol > M ? ld_cg(srow + 2 * M).x : 0.0;
                        } else {
                            if (m < in.ncol) va[u] = ld_cg(srow + 2 * m);
                            if (m != M / 2 && M - m < in.ncol) vb[u] = ld_cg(srow + 2 * (M - m));
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = lane + 32 * (q0 + u);
                if (q0 + u < ITER && it < NF * NPM) {
                    const int f = (NF == 1) ? 0 : it / NPM, m = (NF == 1) ? it : it % NPM;
                    if (m == 0) {
                        A[f * SL] = make_double2(va[u].x + va[u].y, va[u].x - va[u].y);
                    } else if (m == M / 2) {                   // Z' = 2 conj(X_{M/2})
                        A[f * SL + spad(M / 2)] = make_double2(2.0 * va[u].x, -2.0 * va[u].y);
                    } else {
                        const cplx a = va[u], b = make_double2(vb[u].x, -vb[u].y);
                        const cplx e = cadd(a, b);
                        const cplx o = cmulc(csub(a, b), Wh[m]);
                        A[f * SL + spad(m)] = make_double2(e.x - o.y, e.y + o.x);          // e + i o
                        A[f * SL + spad(M - m)] = make_double2(e.x + o.y, -e.y + o.x);     // conj(e) + i conj(o)
                    }
                }
            }
        }
        LG_SYNCWARP();
        prefetch(work + wstride);                            // staging consumed: fetch this warp's next row
        if constexpr (EpiBatched<Epi>::value) {
            fft.template run<true, false>(A, B, W, lane, [](int, int) { return make_double2(0.0, 0.0); },
                                          XinvBatchSt<Epi, WRows<NF>>{epi, rows, fld});
        } else {
            fft.template run<true, false>(A, B, W, lane,
                [](int, int) { return make_double2(0.0, 0.0); },
                [&](int f, int i, cplx v) {
                    const int k = rows.k(f);
                    if (k < 0) return;
                    epi.store(fld, k, rows.y(f), i, v);
                });
        }
        if (lane < NF) {
            const int k = rows.k(lane);
            if (k >= 0) epi.finish_row(fld, k, rows.y(lane));
        }
        LG_SYNCWARP();
    }
}

}  // namespace lg
