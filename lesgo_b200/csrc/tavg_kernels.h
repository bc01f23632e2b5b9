// tavg_kernels.h -- running time averages from the resident fields:
// time_average.f90:176-320 (tavg%compute).  Two kernels: the grid-to-grid interpolations the reference
// takes from functions.f90:51-141 (whose results cross slab seams and are exchanged like there), then ONE
// pass that updates all 26 accumulators.
#pragma once
#include "ops.h"

namespace lg {

enum TavgId {
    TA_U = 0, TA_V, TA_W, TA_W_UV, TA_U_W, TA_V_W, TA_U2, TA_V2, TA_W2, TA_UV, TA_UW, TA_VW, TA_TXX, TA_TYY, TA_TZZ,
    TA_TXY, TA_TXZ, TA_TYZ, TA_P, TA_FX, TA_FY, TA_FZ, TA_CS_OPT2, TA_VORTX, TA_VORTY, TA_VORTZ, TA_N
};

struct TavgInterpArgs {
    const double *u, *v, *w, *dvdx, *dudy, *fza;     // fza may be null
    double *w_uv, *u_w, *v_w, *vortz, *fza_uv;       // outputs, planes 1..nz (see k_tavg_interp)
    int nz, top;
};
// interp_to_uv_grid: out(k) = (in(k+1) + in(k)) / 2 for k = 1..nz-1, out(nz) = out(nz-1) on the top rank;
// interp_to_w_grid:  out(k) = (in(k-1) + in(k)) / 2 for k = 1..nz.   Ghost planes by exchange afterwards.
static __global__ void k_tavg_interp(TavgInterpArgs a, Lay lay, int nx, int ny) {
    const long n = long(nx) * ny * a.nz;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = 1 + int(r / ny);
        const long o = lay.at(k, y, i), om = o - lay.plane, op = o + lay.plane;
        a.u_w[o] = dmul(0.5, dadd(a.u[om], a.u[o]));
        a.v_w[o] = dmul(0.5, dadd(a.v[om], a.v[o]));
        a.vortz[o] = dmul(0.5, dadd(dsub(a.dvdx[om], a.dudy[om]), dsub(a.dvdx[o], a.dudy[o])));
        if (k < a.nz) {
            a.w_uv[o] = dmul(0.5, dadd(a.w[op], a.w[o]));
            if (a.fza) a.fza_uv[o] = dmul(0.5, dadd(a.fza[op], a.fza[o]));
        } else if (a.top) {
            a.w_uv[o] = dmul(0.5, dadd(a.w[o], a.w[om]));
            if (a.fza) a.fza_uv[o] = dmul(0.5, dadd(a.fza[o], a.fza[om]));
        }
    }
}

struct TavgArgs {
    const double *u, *v, *w, *p, *txx, *tyy, *tzz, *txy, *txz, *tyz, *dudz, *dvdz, *dwdx, *dwdy, *cs, *fxa, *fya;
    const double *w_uv, *u_w, *v_w, *vortz, *fza_uv;
    double* acc[TA_N];
    double dt;
    int nz, bottom, top, lbc_mom, ubc_mom, forces;
};
LG_D void tavg_add(double* acc, long o, double v, double dt) { acc[o] = dadd(acc[o], dmul(v, dt)); }

static __global__ void k_tavg_accumulate(TavgArgs a, Lay lay, int nx, int ny) {
    const long n = long(nx) * ny * (a.nz + 1);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = int(r / ny);
        const long o = lay.at(k, y, i);
        const double u = a.u[o], v = a.v[o], w = a.w[o], w_uv = a.w_uv[o];
        double u_w = a.u_w[o], v_w = a.v_w[o], vortz = a.vortz[o];
        if (a.bottom && k == 1) {                                   // time_average.f90:218-227
            vortz = 0.0;
            if (a.lbc_mom > 0) { u_w = 0.0; v_w = 0.0; }
        }
        if (a.top && k == a.nz && a.ubc_mom > 0) { u_w = 0.0; v_w = 0.0; }
        const double dt = a.dt;
        tavg_add(a.acc[TA_U], o, u, dt); tavg_add(a.acc[TA_V], o, v, dt); tavg_add(a.acc[TA_W], o, w, dt);      // :229-234
        tavg_add(a.acc[TA_W_UV], o, w_uv, dt); tavg_add(a.acc[TA_U_W], o, u_w, dt); tavg_add(a.acc[TA_V_W], o, v_w, dt);
        tavg_add(a.acc[TA_U2], o, dmul(u, u), dt); tavg_add(a.acc[TA_V2], o, dmul(v, v), dt);                      // :236-241
        tavg_add(a.acc[TA_W2], o, dmul(w, w), dt); tavg_add(a.acc[TA_UV], o, dmul(u, v), dt);
        tavg_add(a.acc[TA_UW], o, dmul(u_w, w), dt); tavg_add(a.acc[TA_VW], o, dmul(v_w, w), dt);
        tavg_add(a.acc[TA_TXX], o, a.txx[o], dt); tavg_add(a.acc[TA_TYY], o, a.tyy[o], dt);                        // :243-248
        tavg_add(a.acc[TA_TZZ], o, a.tzz[o], dt); tavg_add(a.acc[TA_TXY], o, a.txy[o], dt);
        tavg_add(a.acc[TA_TXZ], o, a.txz[o], dt); tavg_add(a.acc[TA_TYZ], o, a.tyz[o], dt);
        // :204-208 real pressure
        const double ke = dadd(dadd(dmul(u, u), dmul(w_uv, w_uv)), dmul(v, v));
        tavg_add(a.acc[TA_P], o, dsub(a.p[o], dmul(0.5, ke)), dt);
        if (k >= 1) {
            if (a.forces) {                                         // :252-256
                tavg_add(a.acc[TA_FX], o, a.fxa[o], dt); tavg_add(a.acc[TA_FY], o, a.fya[o], dt);
                tavg_add(a.acc[TA_FZ], o, a.fza_uv[o], dt);
            }
            tavg_add(a.acc[TA_CS_OPT2], o, a.cs[o], dt);             // :258
        }
        tavg_add(a.acc[TA_VORTX], o, dsub(a.dwdy[o], a.dvdz[o]), dt);                                               // :260-262
        tavg_add(a.acc[TA_VORTY], o, dsub(a.dudz[o], a.dwdx[o]), dt);
        tavg_add(a.acc[TA_VORTZ], o, vortz, dt);
    }
}

}  // namespace lg
