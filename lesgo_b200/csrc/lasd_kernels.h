// lasd_kernels.h -- pointwise kernels of the Lagrangian scale-dependent dynamic model
// (SURVEY 8(f)-2): lagrange_Sdep.f90:22-430, interpolag_Sdep.f90:21-268 and
// functions.f90:188-265,349-454 (cell_indx_w, trilinear_interp_w).  The 42 test filters per
// plane run through the ordinary x / y FFT passes (lesgo_gpu.cu: filter_fields); everything
// between them is here.  Expression order follows the reference (no FMA contraction), so the
// results agree with the CPU restatement to round-off.
#pragma once
#include "ops.h"

namespace lg {

struct LasdGeom {
    int nx, ny, nz;              // per-rank nz
    int coord, nproc;
    int bottom, top, lbc_mom, ubc_mom;
    double dx, dy, dz, L_x, L_y, L_z;
};

// tensor index order everywhere: 11, 12, 13, 22, 23, 33
LG_D double lasd_contract(const double* a, const double* b) {
    const double d = dadd(dadd(dmul(a[0], b[0]), dmul(a[3], b[3])), dmul(a[5], b[5]));
    const double o = dadd(dadd(dmul(a[1], b[1]), dmul(a[2], b[2])), dmul(a[4], b[4]));
    return dadd(d, dmul(2.0, o));
}
LG_D double lasd_mag(const double* a) { return sqrt(dmul(2.0, lasd_contract(a, a))); }

// ---- step 1 (lagrange_Sdep.f90:87-132): u, v, w on w nodes and their six products -------------------
struct LasdPrepArgs {
    const double *u, *v, *w;     // resident (ld, ny, 0:nz)
    double* A[9];                // ub, vb, wb, ub*ub, ub*vb, ub*wb, vb*vb, vb*wb, wb*wb (addressed by absolute k)
};
static __global__ void k_lasd_prep(LasdPrepArgs a, LasdGeom g, Lay lay, int k0, int k1) {
    const long n = long(g.nx) * g.ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % g.nx);
        long r = t / g.nx;
        const int y = int(r % g.ny), k = k0 + int(r / g.ny);
        const long o = lay.at(k, y, i);
        double ub, vb, wb;
        if (g.bottom && k == 1) {
            ub = a.u[o]; vb = a.v[o];
            wb = g.lbc_mom == 0 ? 0.0 : dmul(0.25, a.w[lay.at(2, y, i)]);
        } else if (g.top && k == g.nz) {
            const long om = lay.at(g.nz - 1, y, i);
            ub = a.u[om]; vb = a.v[om];
            wb = g.ubc_mom == 0 ? 0.0 : dmul(0.25, a.w[om]);
        } else {
            const long om = lay.at(k - 1, y, i);
            ub = dmul(0.5, dadd(a.u[o], a.u[om]));
            vb = dmul(0.5, dadd(a.v[o], a.v[om]));
            wb = a.w[o];
        }
        a.A[0][o] = ub; a.A[1][o] = vb; a.A[2][o] = wb;
        a.A[3][o] = dmul(ub, ub); a.A[4][o] = dmul(ub, vb); a.A[5][o] = dmul(ub, wb);
        a.A[6][o] = dmul(vb, vb); a.A[7][o] = dmul(vb, wb); a.A[8][o] = dmul(wb, wb);
    }
}

// ---- step 2 (:213-218): |S| Sij ------------------------------------------------------------------------
struct LasdSSArgs {
    const double* S[6];
    double* SS[6];
};
static __global__ void k_lasd_ss(LasdSSArgs a, Lay lay, int nx, int ny, int k0, int k1) {
    const long n = long(nx) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, i);
        double s[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = a.S[q][o];
        const double m = lasd_mag(s);
#pragma unroll
        for (int q = 0; q < 6; ++q) a.SS[q][o] = dmul(m, s[q]);
    }
}

// ---- step 3 (:135-168, 205-410): Lij, Qij, Mij, Nij, running averages, Cs_opt2 ---------------------------
struct LasdFinalArgs {
    const double* Tb[9];         // test-filtered      ub, vb, wb, six products
    const double* Th[9];         // test-test-filtered (same order)
    const double* Sb[6];         // test-filtered Sij
    const double* Sh[6];
    const double* SSb[6];        // test-filtered |S| Sij
    const double* SSh[6];
    double *F_LM, *F_MM, *F_QN, *F_NN, *Cs;   // resident
    double delta, lagran_dt, beta_exp;
    int init_F;
};
LG_D void lasd_average(double& F_a, double& F_b, double inst_a, double inst_b, double opftdelta, double lagran_dt) {
    const double zero = 1.0e-24;
    double Tn = fmax(dmul(F_a, F_b), zero);                        // :306-310
    Tn = dmul(opftdelta, pow(Tn, -0.125));
    Tn = fmax(zero, Tn);
    const double dumfac = ddiv(lagran_dt, Tn);                     // :313-314
    const double epsi = ddiv(dumfac, dadd(1.0, dumfac));
    const double om = dsub(1.0, epsi);
    F_a = dadd(dmul(epsi, inst_a), dmul(om, F_a));                 // :316-319
    F_b = dadd(dmul(epsi, inst_b), dmul(om, F_b));
    F_a = fmax(zero, F_a);
}
static __global__ void k_lasd_final(LasdFinalArgs a, LasdGeom g, Lay lay, int k0, int k1) {
    const double zero = 1.0e-24;
    const int ldw = lay.row;
    const long n = long(ldw) * g.ny * (k1 - k0);
    const double cst = dmul(2.0, dmul(a.delta, a.delta));          // const = 2 delta**2
    const double opftdelta = dmul(1.5, a.delta);                   // opftime * delta
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % ldw);
        long r = t / ldw;
        const int y = int(r % g.ny), k = k0 + int(r / g.ny);
        const long o = lay.at(k, y, i);
        if (i >= g.nx) {                                           // :277-278, 325-326, 406-407
            if (a.init_F) { a.F_LM[o] = 1.0; a.F_MM[o] = 1.0; a.F_QN[o] = 1.0; a.F_NN[o] = 1.0; }
            a.Cs[o] = zero;
            continue;
        }
        double L[6], Q[6], M[6], N[6], sb[6], sh[6];
        {
            const double ub = a.Tb[0][o], vb = a.Tb[1][o], wb = a.Tb[2][o];
            L[0] = dsub(a.Tb[3][o], dmul(ub, ub)); L[1] = dsub(a.Tb[4][o], dmul(ub, vb)); L[2] = dsub(a.Tb[5][o], dmul(ub, wb));
            L[3] = dsub(a.Tb[6][o], dmul(vb, vb)); L[4] = dsub(a.Tb[7][o], dmul(vb, wb)); L[5] = dsub(a.Tb[8][o], dmul(wb, wb));
            const double uh = a.Th[0][o], vh = a.Th[1][o], wh = a.Th[2][o];
            Q[0] = dsub(a.Th[3][o], dmul(uh, uh)); Q[1] = dsub(a.Th[4][o], dmul(uh, vh)); Q[2] = dsub(a.Th[5][o], dmul(uh, wh));
            Q[3] = dsub(a.Th[6][o], dmul(vh, vh)); Q[4] = dsub(a.Th[7][o], dmul(vh, wh)); Q[5] = dsub(a.Th[8][o], dmul(wh, wh));
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) { sb[q] = a.Sb[q][o]; sh[q] = a.Sh[q][o]; }
        const double fb = dmul(4.0, lasd_mag(sb));                 // tf1**2 |S_bar|
        const double fh = dmul(16.0, lasd_mag(sh));                // tf2**2 |S_hat|
#pragma unroll
        for (int q = 0; q < 6; ++q) {                              // :243-255
            M[q] = dmul(cst, dsub(a.SSb[q][o], dmul(fb, sb[q])));
            N[q] = dmul(cst, dsub(a.SSh[q][o], dmul(fh, sh[q])));
        }
        const double LM = lasd_contract(L, M), MM = lasd_contract(M, M);      // :258-261
        const double QN = lasd_contract(Q, N), NN = lasd_contract(N, N);
        double F_LM = a.F_LM[o], F_MM = a.F_MM[o], F_QN = a.F_QN[o], F_NN = a.F_NN[o];
        if (a.init_F) { F_MM = MM; F_LM = dmul(0.03, MM); }        // :270-281
        lasd_average(F_LM, F_MM, LM, MM, opftdelta, a.lagran_dt);
        double Cs2 = fmax(zero, ddiv(F_LM, dadd(F_MM, zero)));      // :323-329
        if (a.init_F) { F_NN = NN; F_QN = dmul(0.03, NN); }        // :320-331
        lasd_average(F_QN, F_NN, QN, NN, opftdelta, a.lagran_dt);
        const double Cs4 = fmax(zero, ddiv(F_QN, dadd(F_NN, zero)));          // :376-382
        double Beta = pow(ddiv(Cs4, Cs2), a.beta_exp);                         // :383-384
        if ((g.top && k == g.nz && g.ubc_mom == 0) || (g.bottom && k == 1 && g.lbc_mom == 0)) Beta = 1.0;   // :386-397
        const double Betaclip = fmax(Beta, 0.125);                 // 1 / (tf1 tf2)
        a.F_LM[o] = F_LM; a.F_MM[o] = F_MM; a.F_QN[o] = F_QN; a.F_NN[o] = F_NN;
        a.Cs[o] = fmax(zero, ddiv(Cs2, Betaclip));                  // :400-410
    }
}

// ---- Nu_t = |S| Cs_opt2 l**2 with the coefficient FIELD (sgs_stag_util.f90:221-231) ------------------------
static __global__ void k_nut_field(LasdSSArgs a, const double* __restrict__ Cs, const double* __restrict__ lsq,
                                   double* __restrict__ Nu_t, Lay lay, int nx, int ny, int k0, int k1) {
    const long n = long(nx) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, i);
        double s[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = a.S[q][o];
        Nu_t[o] = dmul(dmul(lasd_mag(s), Cs[o]), lsq[k]);
    }
}

// ---- interpolag_Sdep.f90:21-268 with trilinear_interp_w (functions.f90:349-454) -------------------------------
// cell_indx_w cases 'i' / 'j' (functions.f90:226-252): wraps px in place, returns the 1-based cell index
LG_D int lasd_cell_xy(double& px, double L, double d, int n) {
    double m = fmod(px, L);
    if (m != 0.0 && m < 0.0) m = dadd(m, L);                       // Fortran modulo: result has the sign of L
    px = m;
    if (ddiv(fabs(px), L) < 1.0e-9) return 1;
    if (ddiv(fabs(dsub(px, L)), L) < 1.0e-9) return n;
    return int(floor(ddiv(px, d))) + 1;
}
struct LasdInterpArgs {
    const double *u, *v, *w;
    const double* T[4];          // copies of F_LM, F_MM, F_QN, F_NN (interpolag_Sdep.f90:69-72)
    double* F[4];
    double lagran_dt;
};
// local grid, grid.f90:80-96: z(k) = (coord (nz-1) + k - 1/2) dz, zw = z - dz/2
LG_D double lasd_z(const LasdGeom& g, int k) { return dmul(double(g.coord * (g.nz - 1) + k) - 0.5, g.dz); }
LG_D double lasd_zw(const LasdGeom& g, int k) { return dsub(lasd_z(g, k), g.dz / 2.0); }

#ifndef LG_LASD_EXACT_DIV
#define LG_LASD_EXACT_DIV 0
#endif
static __global__ void k_interpolag(LasdInterpArgs a, LasdGeom g, Lay lay, int k0, int k1) {
    const long n = long(g.nx) * g.ny * (k1 - k0);
    const int nz = g.nz;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % g.nx);
        long r = t / g.nx;
        const int y = int(r % g.ny), k = k0 + int(r / g.ny);
        const long o = lay.at(k, y, i);
        const double xg = dmul(double(i), g.dx), yg = dmul(double(y), g.dy);
        double x0, y0, z0;
        if (g.bottom && k == 1) {                                  // :84-149
            x0 = dsub(xg, dmul(a.u[o], a.lagran_dt));
            y0 = dsub(yg, dmul(a.v[o], a.lagran_dt));
            z0 = g.lbc_mom == 0 ? lasd_zw(g, 1)
                                : dsub(lasd_z(g, 1), dmul(dmul(0.25, a.w[lay.at(2, y, i)]), a.lagran_dt));
        } else if (g.top && k == nz) {                             // :180-241
            const long om = lay.at(nz - 1, y, i);
            x0 = dsub(xg, dmul(a.u[om], a.lagran_dt));
            y0 = dsub(yg, dmul(a.v[om], a.lagran_dt));
            z0 = g.ubc_mom == 0 ? lasd_zw(g, nz)
                                : dsub(lasd_z(g, nz - 1), dmul(dmul(0.25, a.w[om]), a.lagran_dt));
        } else {                                                   // :156-178
            const long om = lay.at(k - 1, y, i);
            x0 = dsub(xg, dmul(dmul(0.5, dadd(a.u[om], a.u[o])), a.lagran_dt));
            y0 = dsub(yg, dmul(dmul(0.5, dadd(a.v[om], a.v[o])), a.lagran_dt));
            z0 = dsub(lasd_zw(g, k), dmul(a.w[o], a.lagran_dt));
        }
        // trilinear_interp_w
        const int ist = lasd_cell_xy(x0, g.L_x, g.dx, g.nx);
        const int jst = lasd_cell_xy(y0, g.L_y, g.dy, g.ny);
        const int ist1 = ist + 1 > g.nx ? 1 : ist + 1;             // autowrap_i, grid.f90:98-104
        const int jst1 = jst + 1 > g.ny ? 1 : jst + 1;
        const double xdiff = dsub(x0, dmul(double(ist - 1), g.dx));
        const double ydiff = dsub(y0, dmul(double(jst - 1), g.dy));
        int kst, kst1;
        double zdiff;
        if (g.bottom && g.lbc_mom > 0 && z0 < lasd_zw(g, 2)) {     // functions.f90:404-418
            kst = 1;
            if (z0 < lasd_z(g, 1)) { kst1 = 1; zdiff = 0.0; }
            else { kst1 = 2; zdiff = dmul(2.0, dsub(z0, lasd_z(g, 1))); }
        } else if (g.top && g.ubc_mom > 0 && z0 > lasd_zw(g, nz - 1)) {   // :419-433
            kst1 = nz;
            if (z0 > lasd_z(g, nz - 1)) { kst = nz; zdiff = 0.0; }
            else { kst = nz - 1; zdiff = dmul(2.0, dsub(z0, lasd_zw(g, nz - 1))); }
        } else {                                                   // :434-443, cell_indx_w 'k' :254-260
            if (ddiv(fabs(dsub(z0, lasd_zw(g, nz))), g.L_z) < 1.0e-9) kst = nz - 1;
            else kst = int(floor(ddiv(dsub(z0, lasd_zw(g, 1)), g.dz))) + 1;
            kst = kst < 0 ? 0 : (kst > nz - 1 ? nz - 1 : kst);     // (a Lagrangian CFL above 1 is outside the reference's assumptions)
            kst1 = kst + 1;
            zdiff = dsub(z0, lasd_zw(g, kst));
        }
        const long o00a = lay.at(kst, jst - 1, ist - 1), o10a = lay.at(kst, jst - 1, ist1 - 1);
        const long o01a = lay.at(kst, jst1 - 1, ist - 1), o11a = lay.at(kst, jst1 - 1, ist1 - 1);
        const long dk = long(kst1 - kst) * lay.plane;
#if LG_LASD_EXACT_DIV
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const double* v = a.T[f];
            const double u1 = dadd(v[o00a], ddiv(dmul(xdiff, dsub(v[o10a], v[o00a])), g.dx));
            const double u2 = dadd(v[o01a], ddiv(dmul(xdiff, dsub(v[o11a], v[o01a])), g.dx));
            const double u3 = dadd(v[o00a + dk], ddiv(dmul(xdiff, dsub(v[o10a + dk], v[o00a + dk])), g.dx));
            const double u4 = dadd(v[o01a + dk], ddiv(dmul(xdiff, dsub(v[o11a + dk], v[o01a + dk])), g.dx));
            const double u5 = dadd(u1, ddiv(dmul(ydiff, dsub(u2, u1)), g.dy));
            const double u6 = dadd(u3, ddiv(dmul(ydiff, dsub(u4, u3)), g.dy));
            a.F[f][o] = dadd(u5, ddiv(dmul(zdiff, dsub(u6, u5)), g.dz));
        }
#else
        // The reference divides each of the 28 increments by dx, dy or dz (functions.f90:445-452); here the three
        // weights are divided once and multiplied in: the same value to within one rounding of each increment, and
        // 3 FP64 divisions per point instead of 28 (the kernel is bound by them: 6.8 ms per call at 512 x 512 x 256).
        const double wx = ddiv(xdiff, g.dx), wy = ddiv(ydiff, g.dy), wz = ddiv(zdiff, g.dz);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const double* v = a.T[f];
            const double u1 = dadd(v[o00a], dmul(wx, dsub(v[o10a], v[o00a])));
            const double u2 = dadd(v[o01a], dmul(wx, dsub(v[o11a], v[o01a])));
            const double u3 = dadd(v[o00a + dk], dmul(wx, dsub(v[o10a + dk], v[o00a + dk])));
            const double u4 = dadd(v[o01a + dk], dmul(wx, dsub(v[o11a + dk], v[o01a + dk])));
            const double u5 = dadd(u1, dmul(wy, dsub(u2, u1)));
            const double u6 = dadd(u3, dmul(wy, dsub(u4, u3)));
            a.F[f][o] = dadd(u5, dmul(wz, dsub(u6, u5)));
        }
#endif
    }
}

}  // namespace lg
