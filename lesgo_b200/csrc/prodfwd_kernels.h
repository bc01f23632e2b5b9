// prodfwd_kernels.h -- the u x omega products and their forward x transform on the 3/2 grid
// (convec.f90:172-305, 207/249/309), marching up z.
//
// Replaces k_xfwd<3nx/2, ProConvec>, whose prologue fetched up to six operands per output value
// from three z planes for each of the three products separately (18 loads per point, 17 words
// per point of DRAM + L2 traffic, latency-bound at 4.2 ms).  Here a block owns one row of the
// 3/2 grid and walks up a chunk of z planes: per plane it reads the six physical rows ONCE
// (prefetched with cp.async while the previous plane is being transformed), keeps what the next
// plane needs -- u(k-1), v(k-1) and the partial sums of cx, cy -- in REGISTERS (each thread owns
// the same row elements at every plane), forms cx(k-1), cy(k-1), cz(k) and transforms the three
// rows as one tile.
#pragma once
#include "ops.h"

namespace lg {

struct ProdArgs {
    const double* src[6];   // 3/2-grid physical fields u, v, w, o1, o2, o3 (planes 0..nz)
    double* dst[3];         // x spectra on the big-y grid of cx, cy, cz (planes 1..nz-1 written)
    long splane, dplane;    // doubles between planes
    int srow, drow;         // doubles between rows
    int ny2, nz;
    int bottom, top, jzLo;
    int chunk, nchunks;
    double scale;           // 1/(nx2*ny2), convec.f90:172
};

// LG_PROD2 = 1: two-stage row transform with LG_PROD2_THR threads.  Measured SLOWER on B200 (512 x 512 x 256: 5.92 / 5.03 /
// 4.30 ms with 64 / 96 / 128 threads against 3.76 ms for the three-stage version with 160 threads): this kernel lives
// on thread-level parallelism across its barriers, profiles/r4_experiments.md.  Off.
#ifndef LG_PROD2
#define LG_PROD2 0
#endif
#ifndef LG_PROD2_THR
#define LG_PROD2_THR 96
#endif
// LG_PROD_TMA = 1: the six physical rows of a plane are fetched by bulk asynchronous copies (cp.async.bulk, one 6 KB
// row each, issued by one thread, completion counted on an mbarrier) instead of 14 cp.async per thread.
#ifndef LG_PROD_TMA
#define LG_PROD_TMA 1
#endif
#ifdef LESGO_EMUL
#undef LG_PROD_TMA
#define LG_PROD_TMA 0
#endif
template <int NX2> struct ProdCfg {
    static constexpr int M = NX2 / 2;
    static constexpr int NC = NX2 / 3;           // nx/2 spectral columns kept
    typedef TileGeom<M> G;
    // two-stage row transform (fft_core.h fft_tile2: radix 16 x 24 for the 3/2 grid of 512) with few, fat threads --
    // the three-stage version issues ~380 instructions per real point and waits on ~9 barriers per plane
    static constexpr bool USE2 = LG_PROD2 && Plan2<M>::on && Plan2<M>::R1 == 16;
    static constexpr int NTHR = USE2 ? LG_PROD2_THR : G::round32(G::threads(3));
    static constexpr int EPT = (M + NTHR - 1) / NTHR;     // row elements (complex pairs) per thread
    static constexpr int SL = SmemLen<M>::value;
    static constexpr int TWL = USE2 ? Plan2Info<M>::twlen : PlanInfo<M>::twlen, NWH = M / 2 + 1;
    static constexpr int TWOFF = USE2 ? PlanInfo<M>::twlen : 0;    // the two-stage rows follow the Stockham tables (lesgo_gpu.cu)
    static constexpr size_t smem = size_t(3 * SL + 6 * M + TWL + NWH) * sizeof(cplx);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int by_regs = 65536 / (NTHR * (USE2 ? 200 : 104)) < 1 ? 1 : 65536 / (NTHR * (USE2 ? 200 : 104));
    static constexpr int MINB0 = by_smem < by_regs ? by_smem : by_regs;
    static constexpr int MINB = MINB0 > 4 ? 4 : MINB0;
};

template <int NX2>
__global__ void __launch_bounds__(ProdCfg<NX2>::NTHR, ProdCfg<NX2>::MINB)
k_prodfwd(const __grid_constant__ ProdArgs a, const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef ProdCfg<NX2> C;
    constexpr int M = C::M, NC = C::NC, SL = C::SL, NTHR = C::NTHR, EPT = C::EPT;
    LG_DYN_SMEM(cplx, sm);
    cplx* rows = sm;                             // cx, cy, cz (padded, natural order)
    cplx* stg = sm + 3 * SL;                     // staging: six physical rows of M complex pairs
    cplx* W = stg + 6 * M;
    cplx* Wh = W + C::TWL;
    load_table(W, Wg + C::TWOFF, C::TWL);
    load_table(Wh, Whg, C::NWH);
    __shared__ double* s_drow[3];
#if LG_PROD_TMA
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    unsigned phase = 0;
    bool pending = false;
#endif
    __syncthreads();

    const int nwork = a.ny2 * a.nchunks;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
        const int y = work % a.ny2, ch = work / a.ny2;
        const int ka = 1 + ch * a.chunk;
        const int kb = ka + a.chunk < a.nz ? ka + a.chunk : a.nz;
        const bool sbchunk = a.bottom && ka == 1;
        const long syoff = long(y) * a.srow, dyoff = long(y) * a.drow;

        // rows to load for step k (bit r: u v w o1 o2 o3); on the bottom rank the pre-step fetches
        // w(2) into slot 0 when jzLo = 1 (convec.f90:174-177)
        auto mask_of = [&](int k) -> int {
            if (k == ka - 1) return sbchunk ? (a.jzLo == 1 ? 1 : 0) : 3;
            if (k < kb) return 63;
            return (a.top && kb == a.nz) ? 0 : 28;
        };
        auto issue = [&](int k) {
            const int mask = mask_of(k);
            const bool stash = sbchunk && k == ka - 1;
#if LG_PROD_TMA
            pending = mask != 0;
            if (threadIdx.x == 0 && mask != 0) {
                mbar_expect_tx(&mbar, unsigned(__popc(unsigned(mask))) * unsigned(M * sizeof(cplx)));
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    if (!((mask >> r) & 1)) continue;
                    const double* srow = (stash ? a.src[2] + 2 * a.splane : a.src[r] + long(k) * a.splane) + syoff;
                    bulk_g2s(stg + r * M, srow, unsigned(M * sizeof(cplx)), &mbar);
                }
            }
#else
            for (int i = threadIdx.x; i < 6 * M; i += NTHR) {
                const int r = i / M, cidx = i - r * M;
                if (!((mask >> r) & 1)) continue;
                const double* srow = (stash ? a.src[2] + 2 * a.splane : a.src[r] + long(k) * a.splane) + syoff;
                cp_async16(stg + i, srow + 2 * cidx);
            }
            cp_async_commit();
#endif
        };

        cplx up[EPT], vp[EPT], px[EPT], py[EPT];          // carried from plane to plane
#pragma unroll
        for (int e = 0; e < EPT; ++e) up[e] = vp[e] = px[e] = py[e] = make_double2(0.0, 0.0);
        issue(ka - 1);
        bool noA = false;
        for (int k = ka - 1; k <= kb; ++k) {
            const int mask = mask_of(k);
#if LG_PROD_TMA
            if (pending) { mbar_wait(&mbar, phase); phase ^= 1u; }
#else
            cp_async_wait_all();
#endif
            __syncthreads();
            // the three output rows of this step are the same for every thread and item of the untangling loop below:
            // formed once here (the loop read them back per item otherwise, ~35 integer instructions each time)
            if (threadIdx.x < 3 && k >= ka)
                s_drow[threadIdx.x] = a.dst[threadIdx.x] + long(threadIdx.x < 2 ? k - 1 : k) * a.dplane + dyoff;
            const bool main = k >= ka && k < kb;
            const bool out_prev = k > ka;
            const bool sb = a.bottom && k == 1;
            const bool haveA = mask != 0 && k >= ka;
            const double sc = a.scale;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = threadIdx.x + e * NTHR;
                if (M % NTHR != 0 && i >= M) break;
                if (k < ka) {                            // pre-step: only the carries
                    if (mask & 1) up[e] = stg[i];
                    if (mask & 2) vp[e] = stg[M + i];
                    continue;
                }
                const int s = spad(i);
                cplx w = make_double2(0.0, 0.0), o1 = w, o2 = w;
                if (haveA) { w = stg[2 * M + i]; o1 = stg[3 * M + i]; o2 = stg[4 * M + i]; }
                const cplx hA = make_double2(0.5 * (w.x * o2.x), 0.5 * (w.y * o2.y));
                const cplx hB = make_double2(0.5 * (w.x * o1.x), 0.5 * (w.y * o1.y));
                cplx ocx = make_double2(0.0, 0.0), ocy = ocx, ocz = ocx;
                if (out_prev) {
                    if (noA || !haveA) { ocx = make_double2(sc * px[e].x, sc * px[e].y); ocy = make_double2(sc * py[e].x, sc * py[e].y); }
                    else {
                        ocx = make_double2(sc * (px[e].x + hA.x), sc * (px[e].y + hA.y));
                        ocy = make_double2(sc * (py[e].x - hB.x), sc * (py[e].y - hB.y));
                    }
                }
                if (main) {
                    const cplx u = stg[i], v = stg[M + i], o3 = stg[5 * M + i];
                    const cplx t1x = make_double2(v.x * (-o3.x), v.y * (-o3.y));
                    const cplx t1y = make_double2(u.x * o3.x, u.y * o3.y);
                    if (sb) {
                        if (a.jzLo == 1) {               // up holds w(2)
                            px[e] = make_double2(t1x.x + 0.5 * up[e].x * o2.x, t1x.y + 0.5 * up[e].y * o2.y);
                            py[e] = make_double2(t1y.x - 0.5 * up[e].x * o1.x, t1y.y - 0.5 * up[e].y * o1.y);
                        } else {
                            px[e] = t1x;
                            py[e] = t1y;
                        }
                    } else {
                        ocz = make_double2(sc * 0.5 * ((u.x + up[e].x) * (-o2.x) + (v.x + vp[e].x) * o1.x),
                                           sc * 0.5 * ((u.y + up[e].y) * (-o2.y) + (v.y + vp[e].y) * o1.y));
                        px[e] = make_double2(t1x.x + hA.x, t1x.y + hA.y);
                        py[e] = make_double2(t1y.x - hB.x, t1y.y - hB.y);
                    }
                    up[e] = u;
                    vp[e] = v;
                }
                rows[s] = ocx;
                rows[SL + s] = ocy;
                rows[2 * SL + s] = ocz;
            }
            if (k >= ka) noA = sb && a.jzLo == 1;
            __syncthreads();
            if (k < kb) issue(k + 1);                    // staging consumed: prefetch the next plane
            if (k < ka) continue;

            // forward transforms of cx(k-1), cy(k-1), cz(k) and the untangled store of nx/2 columns
            if constexpr (C::USE2)
                fft_tile2<M, false, 3, NTHR, true, true, 1, false>(rows, W, [](int f) { return f * SL; },
                    [&](int f, int i) { return rows[f * SL + spad(i)]; },
                    [&](int f, int i, cplx v) { rows[f * SL + spad(i)] = v; });
            else
            fft_tile<M, false, 3, false, NTHR, true, true, 1>(rows, W, [](int f) { return f * SL; },
                [&](int f, int i) { return rows[f * SL + spad(i)]; },
                [&](int f, int i, cplx v) { rows[f * SL + spad(i)] = v; });
            constexpr int NPM = M / 2 + 1;
            for (int it = threadIdx.x; it < 3 * NPM; it += NTHR) {
                const int f = it / NPM, m = it - f * NPM;
                if (f < 2 ? !out_prev : !main) continue;
                const cplx* X = rows + f * SL;
                double* drow = s_drow[f];
                const cplx za = X[spad(m)];
                if (m == 0) {
                    *reinterpret_cast<cplx*>(drow) = make_double2(za.x + za.y, 0.0);
                } else if (m == M / 2) {
                    if (M / 2 < NC) *reinterpret_cast<cplx*>(drow + M) = make_double2(za.x, -za.y);
                } else {
                    const cplx bz = X[spad(M - m)];
                    const cplx b = make_double2(bz.x, -bz.y);
                    const cplx e2 = make_double2(0.5 * (za.x + b.x), 0.5 * (za.y + b.y));
                    const cplx d = make_double2(0.5 * (za.x - b.x), 0.5 * (za.y - b.y));
                    const cplx o = make_double2(d.y, -d.x);
                    const cplx t = cmul(o, Wh[m]);
                    if (m < NC) *reinterpret_cast<cplx*>(drow + 2 * m) = cadd(e2, t);
                    if (M - m < NC) *reinterpret_cast<cplx*>(drow + 2 * (M - m)) = make_double2(e2.x - t.x, -(e2.y - t.y));
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace lg
