// wfft2_kernels.h -- x inverse (half spectrum rows -> real rows) at warp scope with a TWO-STAGE transform.
//
// Same contract as k_xinv / k_xinv_w (XiSrc source descriptor, EpiStore / EpiFused epilogues), other execution
// model: a half-warp owns a row.  The row's M = nx/2 complex points make ONE round trip through the warp's
// private shared-memory slice instead of two, and the real<->half-complex "tangle" step is folded into the loads
// of the first stage instead of being a pass of its own:
//
//     staged spectral row  --(tangle on load)-->  radix-R1 butterflies in registers  -->  shared memory
//                          -->  radix-R2 butterflies (twiddles W_M^{j r} in registers)  -->  epilogue -> global
//
// i.e. per complex point 1 cp.async store + 2 staged loads + 1 store + 1 load through the shared-memory pipe,
// against 1 + 1.5 + 1 + 5 for the three-stage in-place version (k_xinv_w<.., PREF>), which ncu shows at 69-76 %
// of that pipe (profiles/r3a_ncu_full_512x512x32.md).  The next pair of rows is prefetched with cp.async while
// the current pair is transformed; the only synchronisation is __syncwarp().
//   M = 256 (nx = 512):  16 x 16        M = 384 (nx = 768, the 3/2 grid of 512):  16 x 24
#pragma once
#include "wfft_kernels.h"

namespace lg {

template <int M> struct XPlan2 { static constexpr bool on = false; static constexpr int R1 = 1, R2 = 1; };
template <> struct XPlan2<256> { static constexpr bool on = true; static constexpr int R1 = 16, R2 = 16; };
template <> struct XPlan2<384> { static constexpr bool on = true; static constexpr int R1 = 16, R2 = 24; };

template <int NX> struct XW2Cfg {
    static constexpr int M = NX / 2;
    typedef XPlan2<M> P;
    static constexpr int R1 = P::R1, R2 = P::R2;
    static constexpr int T1 = M / R1, T2 = M / R2;            // stage-1 / stage-2 butterflies per row
    static_assert(!P::on || (T2 == 16 && T1 == R2), "a half-warp per row");
    static constexpr int NF = 2;                               // rows per warp
    static constexpr int SLOTS = M + M / R1 + 1;               // padded work buffer of a row: slot s at s + s / R1
    static constexpr int STG = M + 1;                          // staged spectral row (columns 0 .. M)
    static constexpr int WBUF = NF * (SLOTS + STG);            // cplx per warp
    static constexpr int NWH = M / 2 + 1;
    static constexpr int NHI = (R2 - 1) / 8;                   // twiddle rows r = 8, 16, ...
    static constexpr int TWL = (7 + NHI) * 16;                 // W_M^{j r}, r = 1..7, 8, 16, ...; j < 16
    static constexpr int WPB = 4;
    static constexpr int NTHR = 32 * WPB;
    static constexpr size_t smem = size_t(WPB * WBUF + NWH) * sizeof(cplx);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int MINB = by_smem < 4 ? by_smem : 4;
};

// Z'_i of the packed half-length inverse transform from the staged half spectrum X (columns >= ncol read as 0:
// the staging buffer is zero there):  E' = X_m + conj(X_{M-m}),  O' = (X_m - conj(X_{M-m})) conj(W_N^m),
// Z'_m = E' + i O',  Z'_{M-m} = conj(E') + i conj(O')
template <int M>
LG_D cplx tangle_at(const cplx* X, const cplx* Wh, int i) {
    if (i == 0) {
        const double x0 = X[0].x, xm = X[M].x;
        return make_double2(x0 + xm, x0 - xm);
    }
    if (i == M / 2) {
        const cplx a = X[M / 2];
        return make_double2(2.0 * a.x, -2.0 * a.y);
    }
    const int m = i < M / 2 ? i : M - i;
    const cplx a = X[m], bz = X[M - m];
    const cplx b = make_double2(bz.x, -bz.y);
    const cplx e = cadd(a, b);
    const cplx o = cmulc(csub(a, b), Wh[m]);
    return i < M / 2 ? make_double2(e.x - o.y, e.y + o.x) : make_double2(e.x + o.y, -e.y + o.x);
}

template <int NX, class Epi>
__global__ void __launch_bounds__(XW2Cfg<NX>::NTHR, XW2Cfg<NX>::MINB)
k_xinv_w2(const __grid_constant__ XiSrc in, const __grid_constant__ Epi epi, int nfields, int ny, int k0, int nplanes,
          const cplx* __restrict__ W2g, const cplx* __restrict__ Whg) {
    typedef XW2Cfg<NX> C;
    constexpr int M = C::M, R1 = C::R1, R2 = C::R2, T1 = C::T1;
    LG_DYN_SMEM(cplx, sm);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int f = lane >> 4, j = lane & 15;                   // row of the pair, butterfly index
    cplx* buf = sm + wib * C::WBUF + f * (C::SLOTS + C::STG);
    cplx* ST = buf + C::SLOTS;
    cplx* Wh = sm + C::WPB * C::WBUF;
    load_table(Wh, Whg, C::NWH);
    for (int m = j; m < C::STG; m += 16) ST[m] = make_double2(0.0, 0.0);     // columns >= ncol stay zero
#if LG_XW_TMA
    __shared__ __align__(8) unsigned long long mbar[C::WPB];
    if (lane == 0) mbar_init(&mbar[wib], 1);
    unsigned phase = 0;
#endif
    // stage-2 twiddles W_M^{j r} depend only on the lane: r = 1..7 and r = 8, 16 in registers for the whole loop
    cplx lo[8], hi[C::NHI + 1];
#pragma unroll
    for (int r = 1; r < 8; ++r) lo[r] = W2g[(r - 1) * 16 + j];
#pragma unroll
    for (int h = 1; h <= C::NHI; ++h) hi[h] = W2g[(6 + h) * 16 + j];
    __syncthreads();

    const unsigned nrows = unsigned(ny) * unsigned(nplanes);
    const unsigned npairs = (nrows + 1) / 2;
    const unsigned nwork = npairs * unsigned(nfields);
    const unsigned wstride = gridDim.x * C::WPB;
    const int nc = in.ncol < M + 1 ? in.ncol : M + 1;
    auto prefetch = [&](unsigned work) {
#if LG_XW_TMA
        // lane 0 fetches both rows of the pair; a pair past the end or an empty row still completes the phase
        if (lane == 0 && work < nwork) {
            const unsigned ra = (work / unsigned(nfields)) * 2;
            const int pf = int(work % unsigned(nfields));
            const unsigned nlive = (ra < nrows ? 1u : 0u) + (ra + 1 < nrows ? 1u : 0u);
            mbar_expect_tx(&mbar[wib], nlive * unsigned(nc) * unsigned(sizeof(cplx)));
            for (unsigned ff = 0; ff < 2 && nc > 0; ++ff) {
                const unsigned r = ra + ff;
                if (r >= nrows) continue;
                const double* srow = in.src[pf] + poff(k0 + int(r / unsigned(ny)), in.plane, in.ring) + long(r % unsigned(ny)) * in.row;
                bulk_g2s(sm + wib * C::WBUF + ff * (C::SLOTS + C::STG) + C::SLOTS, srow, unsigned(nc) * unsigned(sizeof(cplx)), &mbar[wib]);
            }
        }
#else
        if (work < nwork) {
            const unsigned r = (work / unsigned(nfields)) * 2 + f;
            if (r < nrows) {
                const int pf = int(work % unsigned(nfields));
                const double* srow = in.src[pf] + poff(k0 + int(r / unsigned(ny)), in.plane, in.ring) + long(r % unsigned(ny)) * in.row;
                for (int m = j; m < nc; m += 16) cp_async16(ST + m, srow + 2 * m);
            }
        }
        cp_async_commit();
#endif
    };
    prefetch(blockIdx.x * C::WPB + wib);
    for (unsigned work = blockIdx.x * C::WPB + wib; work < nwork; work += wstride) {
        const int fld = int(work % unsigned(nfields));
        const unsigned r0 = (work / unsigned(nfields)) * 2 + f;
        const bool live = r0 < nrows;
        const int k = k0 + int(r0 / unsigned(ny)), y = int(r0 % unsigned(ny));
#if LG_XW_TMA
        mbar_wait(&mbar[wib], phase);
        phase ^= 1u;
#else
        cp_async_wait_all();
#endif
        LG_SYNCWARP();
        // ---- stage 1: radix R1 on Z'_{it + T1 r}, no twiddles; slot R1*it + r (padded: (R1+1)*it + r) --------
#pragma unroll
        for (int q = 0; q < (T1 + 15) / 16; ++q) {
            const int it = j + 16 * q;
            if (T1 % 16 == 0 || it < T1) {
                cplx v[R1];
#pragma unroll
                for (int r = 0; r < R1; ++r) v[r] = cswap(tangle_at<M>(ST, Wh, it + T1 * r));
                Dft<R1>::run(v);
                cplx* p = buf + (R1 + 1) * it;
#pragma unroll
                for (int r = 0; r < R1; ++r) p[r] = cswap(v[r]);
            }
        }
        LG_SYNCWARP();
        prefetch(work + wstride);                              // staging consumed: fetch this warp's next pair
        // ---- stage 2: radix R2, Ns = R1: butterfly j reads slots j + R1 r, writes points j + R1 r ----------------
        {
            cplx v[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) v[r] = buf[j + (R1 + 1) * r];
#pragma unroll
            for (int r = 1; r < R2; ++r) {
                const cplx w = r < 8 ? lo[r] : ((r & 7) == 0 ? hi[r >> 3] : cmul(hi[r >> 3], lo[r & 7]));
                v[r] = cmulc(v[r], w);                          // inverse: conjugated twiddles
            }
#pragma unroll
            for (int r = 0; r < R2; ++r) v[r] = cswap(v[r]);
            Dft<R2>::run(v);
            if (live) {
#pragma unroll
                for (int r = 0; r < R2; ++r) epi.store(fld, k, y, j + R1 * r, cswap(v[r]));
                if (j == 0) epi.finish_row(fld, k, y);
            }
        }
        LG_SYNCWARP();
    }
    cp_async_wait_all();
}

}  // namespace lg
