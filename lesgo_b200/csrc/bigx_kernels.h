// bigx_kernels.h -- the x direction of the dealiased products in ONE kernel (convec.f90:90-92,
// 165-167, 172-305): inverse x transforms of u, v, w, omega_1..3 on the 3/2 grid, the u x omega
// products, and the forward x transforms of the three products, without ever writing a
// 3/2-grid physical field to memory.
//
// Round 1 ran this as k_xinv<3nx/2> (writes six physical fields, 13.5 words per point),
// then k_xfwd<3nx/2, ProConvec> (reads them back with their k+-1 neighbours): 17 GB of the
// step's 107 GB of DRAM traffic and 8.6 of its 29 ms.  Here a block owns one y row of the 3/2
// grid and MARCHES UP z through a chunk of planes.  The products couple plane k to k-1 (u, v in
// cz) and to k+1 (w omega in cx, cy), so the block carries four rows from plane to plane in
// shared memory -- u(k-1), v(k-1) and the partial sums
//     Px(k) = v(k) (-o3(k)) + w(k) o2(k) / 2,      Py(k) = u(k) o3(k) - w(k) o1(k) / 2
// -- and emits cx(k-1), cy(k-1) one step late, when w(k) o(k) is known.  Per plane the block
// reads six spectral rows (prefetched with cp.async while the previous plane is being
// computed) and writes three.
//
// Shared-memory rows (padded, natural order): U[2] V[2] (ping-pong: current / previous plane)
// W O1 O2 O3 PX PY.  The inverse transforms of a plane run as one 6-row fft_tile; the products
// are formed in place (cx -> O3 row, cy -> W row, cz -> O1 row) and transformed as one 3-row tile.
#pragma once
#include "ops.h"

namespace lg {

struct BigxArgs {
    const double* src[6];   // x spectra on the big-y grid of u, v, w, o1, o2, o3 (planes 0..nz)
    double* dst[3];         // x spectra on the big-y grid of cx, cy, cz (planes 1..nz-1 written)
    long plane;             // doubles between planes (src and dst)
    int row;                // doubles between rows
    int ny2, nz;
    int bottom, top, jzLo;
    int chunk, nchunks;     // planes per z chunk, chunks per row
    double scale;           // 1/(nx2*ny2), convec.f90:172
};

template <int NX2> struct BigxCfg {
    static constexpr int M = NX2 / 2;            // half length of a 3/2-grid row
    static constexpr int NC = NX2 / 3;           // nx/2: spectral columns that carry data
    typedef TileGeom<M> G;
    static constexpr int NTHR = G::round32(G::threads(6));
    static constexpr int SL = SmemLen<M>::value;
    static constexpr int NROWS = 10;
    static constexpr int TWL = PlanInfo<M>::twlen, NWH = M / 2 + 1;
    static constexpr size_t smem = size_t(NROWS * SL + 6 * NC + TWL + NWH) * sizeof(cplx);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int by_regs = 65536 / (NTHR * 80) < 1 ? 1 : 65536 / (NTHR * 80);
    static constexpr int MINB = by_smem < by_regs ? by_smem : by_regs;
};

enum { BX_U0 = 0, BX_V0 = 1, BX_U1 = 2, BX_V1 = 3, BX_W = 4, BX_O1 = 5, BX_O2 = 6, BX_O3 = 7, BX_PX = 8, BX_PY = 9 };

template <int NX2>
__global__ void __launch_bounds__(BigxCfg<NX2>::NTHR, BigxCfg<NX2>::MINB)
k_bigx(const __grid_constant__ BigxArgs a, const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef BigxCfg<NX2> C;
    constexpr int M = C::M, NC = C::NC, SL = C::SL, NTHR = C::NTHR;
    LG_DYN_SMEM(cplx, sm);
    cplx* rows = sm;
    cplx* stg = sm + C::NROWS * SL;              // staging: 6 spectral rows of NC columns
    cplx* W = stg + 6 * NC;
    cplx* Wh = W + C::TWL;
    load_table(W, Wg, C::TWL);
    load_table(Wh, Whg, C::NWH);
    __syncthreads();

    const int nwork = a.ny2 * a.nchunks;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
        const int y = work % a.ny2, ch = work / a.ny2;       // blocks resident together share planes
        const int ka = 1 + ch * a.chunk;
        const int kb = ka + a.chunk < a.nz ? ka + a.chunk : a.nz;
        const bool sbchunk = a.bottom && ka == 1;
        const long yoff = long(y) * a.row;

        // rows to load for step k: bit r = tile row r (u v w o1 o2 o3); tile row 0 may be redirected
        auto mask_of = [&](int k) -> int {
            if (k == ka - 1) return sbchunk ? (a.jzLo == 1 ? 1 : 0) : 3;   // w(2) stash, or u(k), v(k)
            if (k < kb) return 63;
            return (a.top && kb == a.nz) ? 0 : 28;                          // w, o1, o2 of plane kb
        };
        auto issue = [&](int k) {
            const int mask = mask_of(k);
            const bool stash = sbchunk && k == ka - 1;                      // tile row 0 <- w(2)
            for (int i = threadIdx.x; i < 6 * NC; i += NTHR) {
                const int r = i / NC, cidx = i - r * NC;
                if (!((mask >> r) & 1)) continue;
                const double* srow = (stash ? a.src[2] + 2 * a.plane : a.src[r] + long(k) * a.plane) + yoff;
                cp_async16(stg + i, srow + 2 * cidx);
            }
            cp_async_commit();
        };

        issue(ka - 1);
        bool noA = false;                         // cx(k-1), cy(k-1) already hold their w*omega term
        for (int k = ka - 1; k <= kb; ++k) {
            const int par = k & 1;
            const int mask = mask_of(k);
            auto rowof = [&](int r) { return (r < 2 ? r + 2 * par : r + 2) * SL; };   // tile row -> shared row
            cp_async_wait_all();
            __syncthreads();
            if (mask) {
                // tangle (see k_xinv): staging -> rows;  columns >= NC are zero, X_M is absent
                constexpr int NPM = M / 2 + 1;
                for (int it = threadIdx.x; it < 6 * NPM; it += NTHR) {
                    const int r = it / NPM, m = it - r * NPM;
                    if (!((mask >> r) & 1)) continue;
                    const cplx* s = stg + r * NC;
                    cplx* d = rows + rowof(r);
                    if (m == 0) {
                        const double x0 = s[0].x;
                        d[0] = make_double2(x0, x0);
                    } else if (m == M / 2) {
                        const cplx x = (M / 2 < NC) ? s[M / 2] : make_double2(0.0, 0.0);
                        d[spad(M / 2)] = make_double2(2.0 * x.x, -2.0 * x.y);
                    } else {
                        const cplx xa = (m < NC) ? s[m] : make_double2(0.0, 0.0);
                        const cplx xb = (M - m < NC) ? s[M - m] : make_double2(0.0, 0.0);
                        const cplx b = make_double2(xb.x, -xb.y);
                        const cplx e = cadd(xa, b);
                        const cplx o = cmulc(csub(xa, b), Wh[m]);
                        d[spad(m)] = make_double2(e.x - o.y, e.y + o.x);
                        d[spad(M - m)] = make_double2(e.x + o.y, -e.y + o.x);
                    }
                }
                __syncthreads();
            }
            if (k < kb) issue(k + 1);             // staging is free: prefetch the next plane
            if (mask) {
                fft_tile<M, true, 6, false, NTHR, true, true, 1>(rows, W, rowof,
                    [&](int f, int i) { return rows[rowof(f) + spad(i)]; },
                    [&](int f, int i, cplx v) { rows[rowof(f) + spad(i)] = v; });
            }
            if (k < ka) continue;

            // ---- products of this step, in place ----
            const bool main = k < kb;
            const bool out_prev = k > ka;                         // cx(k-1), cy(k-1) complete now
            const bool sb = a.bottom && k == 1;
            const cplx* uc = rows + (BX_U0 + 2 * par) * SL;
            const cplx* vc = rows + (BX_V0 + 2 * par) * SL;
            const cplx* up = rows + (BX_U0 + 2 * (par ^ 1)) * SL;   // u(k-1); w(2) on the bottom rank at k = 1
            const cplx* vp = rows + (BX_V0 + 2 * (par ^ 1)) * SL;
            cplx* rw = rows + BX_W * SL;
            cplx* r1 = rows + BX_O1 * SL;
            cplx* r2 = rows + BX_O2 * SL;
            cplx* r3 = rows + BX_O3 * SL;
            cplx* px = rows + BX_PX * SL;
            cplx* py = rows + BX_PY * SL;
            const bool haveA = mask != 0;                         // w, o1, o2 of plane k are in the rows
            const double sc = a.scale;
            for (int i = threadIdx.x; i < M; i += NTHR) {
                const int s = spad(i);
                cplx w = make_double2(0.0, 0.0), o1 = w, o2 = w;
                if (haveA) { w = rw[s]; o1 = r1[s]; o2 = r2[s]; }
                const cplx hA = make_double2(0.5 * (w.x * o2.x), 0.5 * (w.y * o2.y));   // w o2 / 2
                const cplx hB = make_double2(0.5 * (w.x * o1.x), 0.5 * (w.y * o1.y));   // w o1 / 2
                cplx ocx = make_double2(0.0, 0.0), ocy = ocx, ocz = ocx;
                if (out_prev) {
                    const cplx qx = px[s], qy = py[s];
                    if (noA || !haveA) { ocx = make_double2(sc * qx.x, sc * qx.y); ocy = make_double2(sc * qy.x, sc * qy.y); }
                    else {
                        ocx = make_double2(sc * (qx.x + hA.x), sc * (qx.y + hA.y));
                        ocy = make_double2(sc * (qy.x - hB.x), sc * (qy.y - hB.y));
                    }
                }
                if (main) {
                    const cplx u = uc[s], v = vc[s], o3 = r3[s];
                    const cplx t1x = make_double2(v.x * (-o3.x), v.y * (-o3.y));
                    const cplx t1y = make_double2(u.x * o3.x, u.y * o3.y);
                    if (sb) {
                        // bottom wall, convec.f90:174-177, 217-220: w(2) o(jzLo) / 2 replaces the k, k+1 average;
                        // jzLo = 1: w(2) was stashed in the previous-u row;  jzLo = 2: added next step as usual
                        if (a.jzLo == 1) {
                            const cplx w2 = up[s];
                            px[s] = make_double2(t1x.x + 0.5 * w2.x * o2.x, t1x.y + 0.5 * w2.y * o2.y);
                            py[s] = make_double2(t1y.x - 0.5 * w2.x * o1.x, t1y.y - 0.5 * w2.y * o1.y);
                        } else {
                            px[s] = t1x;
                            py[s] = t1y;
                        }
                    } else {
                        const cplx pu = up[s], pv = vp[s];
                        ocz = make_double2(sc * 0.5 * ((u.x + pu.x) * (-o2.x) + (v.x + pv.x) * o1.x),
                                           sc * 0.5 * ((u.y + pu.y) * (-o2.y) + (v.y + pv.y) * o1.y));
                        px[s] = make_double2(t1x.x + hA.x, t1x.y + hA.y);
                        py[s] = make_double2(t1y.x - hB.x, t1y.y - hB.y);
                    }
                }
                r3[s] = ocx;          // tile row 0 of the forward transform
                rw[s] = ocy;          // tile row 1
                r1[s] = ocz;          // tile row 2
            }
            noA = sb && a.jzLo == 1;
            __syncthreads();

            // ---- forward transforms of cx(k-1), cy(k-1), cz(k) and the untangled store ----
            auto frow = [&](int f) { return (f == 0 ? BX_O3 : (f == 1 ? BX_W : BX_O1)) * SL; };
            fft_tile<M, false, 3, false, NTHR, true, true, 1>(rows, W, frow,
                [&](int f, int i) { return rows[frow(f) + spad(i)]; },
                [&](int f, int i, cplx v) { rows[frow(f) + spad(i)] = v; });
            constexpr int NPM = M / 2 + 1;
            for (int it = threadIdx.x; it < 3 * NPM; it += NTHR) {
                const int f = it / NPM, m = it - f * NPM;
                if (f < 2 ? !out_prev : !main) continue;
                const cplx* X = rows + frow(f);
                double* drow = a.dst[f] + long(f < 2 ? k - 1 : k) * a.plane + yoff;
                const cplx za = X[spad(m)];
                if (m == 0) {
                    *reinterpret_cast<cplx*>(drow) = make_double2(za.x + za.y, 0.0);
                } else if (m == M / 2) {
                    if (M / 2 < NC) *reinterpret_cast<cplx*>(drow + M) = make_double2(za.x, -za.y);
                } else {
                    const cplx bz = X[spad(M - m)];
                    const cplx b = make_double2(bz.x, -bz.y);
                    const cplx e = make_double2(0.5 * (za.x + b.x), 0.5 * (za.y + b.y));
                    const cplx d = make_double2(0.5 * (za.x - b.x), 0.5 * (za.y - b.y));
                    const cplx o = make_double2(d.y, -d.x);
                    const cplx t = cmul(o, Wh[m]);
                    if (m < NC) *reinterpret_cast<cplx*>(drow + 2 * m) = cadd(e, t);
                    if (M - m < NC) *reinterpret_cast<cplx*>(drow + 2 * (M - m)) = make_double2(e.x - t.x, -(e.y - t.y));
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace lg
