// ops.h -- prologue / epilogue functors fused into the x passes, and the pointwise
// z-direction kernels (finite differences, tridiagonal sweeps, time-stepping glue).
#pragma once
#include "fft_kernels.h"

namespace lg {

// Products are formed without FMA contraction where the reference's expression order
// matters for parity with the (non-FMA) CPU restatement.
#ifdef LESGO_EMUL
LG_HD double dmul(double a, double b) { return a * b; }
LG_HD double dadd(double a, double b) { return a + b; }
LG_HD double dsub(double a, double b) { return a - b; }
LG_HD double ddiv(double a, double b) { return a / b; }
#else
LG_D double dmul(double a, double b) { return __dmul_rn(a, b); }
LG_D double dadd(double a, double b) { return __dadd_rn(a, b); }
LG_D double dsub(double a, double b) { return __dsub_rn(a, b); }
LG_D double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#endif

struct Lay {          // addressing of a (row, plane)-strided real array
    long plane;       // doubles between z planes
    int row;          // doubles between y rows
    LG_HD long at(int k, int y, int i) const { return long(k) * plane + long(y) * row + i; }
};

LG_HD double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---- x-forward prologues ---------------------------------------------------------
// scale * f   (derivatives.f90:185-190, convec.f90:75-77, press_stag_array.f90:78-80)
struct ProScale {
    const double* src[kMaxFields];
    Lay lay;
    double scale;
    LG_D double2 load(int fld, int k, int y, int j) const {
        double2 v = ld2(src[fld] + lay.at(k, y, 2 * j));
        return make_double2(dmul(scale, v.x), dmul(scale, v.y));
    }
};

// vorticity with the wall-plane special cases (convec.f90:97-153); fields 0,1,2 = RHSx,y,z
struct ProVort {
    const double *dudy, *dudz, *dvdx, *dvdz, *dwdx, *dwdy;
    Lay lay;
    double scale;
    int nz, bottom, top, lbc_mom, ubc_mom;   // bottom = (coord == 0), top = (coord == nproc-1)
    // The case analysis depends on (field, plane) only, so a pass does it once per ROW (row()) and then reads the
    // row's elements with straight-line code (load_row()): with the analysis inside load() every element was a
    // load -> branch -> use chain of its own and a thread's 8 elements cost 8 DRAM round trips instead of one.
    //   kind 0: 0;  1: scale*(a - b);  2: scale*(0.5*(a + b) - c);  3: scale*(c - 0.5*(a + b))
    struct Row { const double *pa, *pb, *pc; int kind; };
    LG_D Row row(int fld, int k, int y) const {          // k < 0: a row past the end, reads as zeros
        Row r;
        r.pa = r.pb = r.pc = nullptr; r.kind = 0;
        if (k < 0) return r;
        const long o = lay.at(k, y, 0);
        if (fld == 2) { r.pa = dvdx + o; r.pb = dudy + o; r.kind = 1; return r; }
        const bool sb = bottom && k == 1;
        const bool st = top && k == nz && ubc_mom > 0;
        if (sb && lbc_mom == 0) return r;
        if (sb || st) {
            // 0.5*(dw(k1)+dw(k2)) -/+ d(u|v)dz(kz)
            const int k1 = sb ? 1 : nz - 1, k2 = sb ? 2 : nz, kz = sb ? 1 : nz - 1;
            const long o1 = lay.at(k1, y, 0), o2 = lay.at(k2, y, 0), oz = lay.at(kz, y, 0);
            if (fld == 0) { r.pa = dwdy + o1; r.pb = dwdy + o2; r.pc = dvdz + oz; r.kind = 2; }
            else { r.pa = dwdx + o1; r.pb = dwdx + o2; r.pc = dudz + oz; r.kind = 3; }
            return r;
        }
        if (fld == 0) { r.pa = dwdy + o; r.pb = dvdz + o; }
        else { r.pa = dudz + o; r.pb = dwdx + o; }
        r.kind = 1;
        return r;
    }
    LG_D double2 load_row(const Row& r, int j) const {
        double2 a = make_double2(0.0, 0.0), b = a, c = a;
        if (r.kind >= 1) { a = ld2(r.pa + 2 * j); b = ld2(r.pb + 2 * j); }
        if (r.kind >= 2) c = ld2(r.pc + 2 * j);
        const double2 h = make_double2(dmul(0.5, dadd(a.x, b.x)), dmul(0.5, dadd(a.y, b.y)));
        const double2 p = r.kind == 1 ? a : (r.kind == 2 ? h : c);
        const double2 q = r.kind == 1 ? b : (r.kind == 2 ? c : h);
        return make_double2(dmul(scale, dsub(p.x, q.x)), dmul(scale, dsub(p.y, q.y)));      // kind 0: scale*(0 - 0)
    }
    LG_D double2 load(int fld, int k, int y, int j) const { return load_row(row(fld, k, y), j); }
};

// u x omega on the 3/2 grid (convec.f90:172-305); fields 0,1,2 = cx, cy, cz
struct ProConvec {
    const double *u, *v, *w, *o1, *o2, *o3;   // *_big arrays, planes 0..nz
    Lay lay;
    double scale;                              // 1/(nx2*ny2)
    int nz, bottom, top, jzLo;
    LG_D double2 load(int fld, int k, int y, int j) const {
        const long o = lay.at(k, y, 2 * j);
        const bool sb = bottom && k == 1;
        const bool st = top && k == nz - 1;
        double2 r;
        if (fld == 2) {
            if (sb || k == nz) return make_double2(0.0, 0.0);
            const long om = lay.at(k - 1, y, 2 * j);
            double2 uu = ld2(u + o), um = ld2(u + om), vv = ld2(v + o), vm = ld2(v + om);
            double2 w1 = ld2(o1 + o), w2 = ld2(o2 + o);
            r.x = dmul(dmul(scale, 0.5), dadd(dmul(dadd(uu.x, um.x), -w2.x), dmul(dadd(vv.x, vm.x), w1.x)));
            r.y = dmul(dmul(scale, 0.5), dadd(dmul(dadd(uu.y, um.y), -w2.y), dmul(dadd(vv.y, vm.y), w1.y)));
            return r;
        }
        if (k == nz) return make_double2(0.0, 0.0);     // plane nz of cx, cy is never valid
        // first term: cx: v*(-o3), cy: u*o3
        double2 f = ld2((fld == 0 ? v : u) + o), z3 = ld2(o3 + o);
        const double sg = fld == 0 ? -1.0 : 1.0;        // sign on o3 (cx) / on o1 (cy) below is opposite
        double2 t1 = make_double2(dmul(f.x, sg * z3.x), dmul(f.y, sg * z3.y));
        const double* oz = fld == 0 ? o2 : o1;          // cx uses +o2, cy uses -o1
        const double so = fld == 0 ? 1.0 : -1.0;
        double2 t2;
        if (st) {
            // top rank, plane nz-1: 0.5*w(nz-1)*o(jzHi = nz-1)   (:186-189, :229-232)
            double2 ww = ld2(w + o), zz = ld2(oz + o);
            t2 = make_double2(dmul(dmul(0.5, ww.x), so * zz.x), dmul(dmul(0.5, ww.y), so * zz.y));
        } else if (sb) {
            // bottom rank, plane 1: 0.5*w(2)*o(jzLo)             (:174-177, :217-220)
            double2 ww = ld2(w + lay.at(2, y, 2 * j)), zz = ld2(oz + lay.at(jzLo, y, 2 * j));
            t2 = make_double2(dmul(dmul(0.5, ww.x), so * zz.x), dmul(dmul(0.5, ww.y), so * zz.y));
        } else {
            const long op = lay.at(k + 1, y, 2 * j);
            double2 wp = ld2(w + op), zp = ld2(oz + op), ww = ld2(w + o), zz = ld2(oz + o);
            t2 = make_double2(dmul(0.5, dadd(dmul(wp.x, so * zp.x), dmul(ww.x, so * zz.x))),
                              dmul(0.5, dadd(dmul(wp.y, so * zp.y), dmul(ww.y, so * zz.y))));
        }
        return make_double2(dmul(scale, dadd(t1.x, t2.x)), dmul(scale, dadd(t1.y, t2.y)));
    }
};

// ---- x-inverse epilogues ------------------------------------------------------------
// plain store (every routine-level entry point)
struct EpiStore {
    double* dst[kMaxFields];
    Lay lay;
    int nx;          // real row length; the two pad reals are zeroed when pad != 0
    int pad;
    LG_D void store(int fld, int k, int y, int j, double2 v) const {
        *reinterpret_cast<double2*>(dst[fld] + lay.at(k, y, 2 * j)) = v;
    }
    LG_D void finish_row(int fld, int k, int y) const {
        if (pad) *reinterpret_cast<double2*>(dst[fld] + lay.at(k, y, nx)) = make_double2(0.0, 0.0);
    }
};

// store + time-stepping glue, used by lesgo_gpu_step only (a separate type so the plain epilogue
// stays lean: the extra state cost 0.5 ms in the 3/2-grid x-inverse when it lived in EpiStore).
// mode 1: the value is the convective term cc; fuses the RHS assembly, the Euler start and the
//         AB2 update (main.f90:211-214, 229-232, 273-280, 287-296):
//             rhs = -cc - divt + force;  [first step: rhs_f = rhs];  u += dt*(tadv1*rhs + tadv2*rhs_f)
//         on planes k <= kmax[fld] (the other planes just get the plain store).
// mode 2: the value is a pressure-gradient component; fuses main.f90:321-326 and project
//         (forcing.f90:171-207):  dpd = value;  rhs -= value;  u += dt*(-tadv1*value)
struct EpiFused {
    double* dst[kMaxFields];
    Lay lay;
    int nx;
    int pad;
    int mode;
    const double* divt[3];
    double* rhs_f[3];
    double* u[3];
    double force[3];
    int kmax[3];
    int first_step;
    double dt, t1, t2;
    const double* fa[3];     // applied body force fxa, fya, fza added on planes k <= kfa (main.f90:264-266), or null
    int kfa;
    // The operands of one output element.  load() and apply() are separate so that a pass can issue the loads of
    // ALL the elements a lane is about to produce before it stores any of them: with load-use-store per element
    // the possibly-aliasing stores serialise the loads, one DRAM round trip per element (8 per row and lane).
    struct Ops { double2 a, b, c, d; };
    LG_D Ops load(int fld, int k, int y, int j) const {
        Ops p;
        p.a = p.b = p.c = p.d = make_double2(0.0, 0.0);
        const long o = lay.at(k, y, 2 * j);
        if (mode == 1 && k <= kmax[fld]) {
            p.a = ld2(divt[fld] + o);
            p.b = ld2(u[fld] + o);
            if (!first_step) p.c = ld2(rhs_f[fld] + o);
            if (fa[fld] && k <= kfa) p.d = ld2(fa[fld] + o);
        } else if (mode == 2) {
            p.a = ld2(rhs_f[fld] + o);                                   // rhs_f[] holds RHS here
            p.b = ld2(u[fld] + o);
        }
        return p;
    }
    LG_D void apply(int fld, int k, int y, int j, double2 v, const Ops& p) const {
        const long o = lay.at(k, y, 2 * j);
        if (mode == 1 && k <= kmax[fld]) {
            const double2 vb = p.a, vu = p.b;
            const double f = force[fld];
            double2 nr = make_double2(dadd(dsub(-v.x, vb.x), f), dadd(dsub(-v.y, vb.y), f));
            if (fa[fld] && k <= kfa) nr = make_double2(dadd(nr.x, p.d.x), dadd(nr.y, p.d.y));
            double2 vf;
            if (first_step) { vf = nr; *reinterpret_cast<double2*>(rhs_f[fld] + o) = nr; }
            else vf = p.c;
            *reinterpret_cast<double2*>(dst[fld] + o) = nr;
            *reinterpret_cast<double2*>(u[fld] + o) =
                make_double2(dadd(vu.x, dmul(dt, dadd(dmul(t1, nr.x), dmul(t2, vf.x)))),
                             dadd(vu.y, dmul(dt, dadd(dmul(t1, nr.y), dmul(t2, vf.y)))));
            return;
        }
        *reinterpret_cast<double2*>(dst[fld] + o) = v;
        if (mode == 2) {
            const double2 vr = p.a, vu = p.b;
            *reinterpret_cast<double2*>(rhs_f[fld] + o) = make_double2(dsub(vr.x, v.x), dsub(vr.y, v.y));
            *reinterpret_cast<double2*>(u[fld] + o) = make_double2(dadd(vu.x, dmul(dt, dmul(-t1, v.x))),
                                                                   dadd(vu.y, dmul(dt, dmul(-t1, v.y))));
        }
    }
    LG_D void store(int fld, int k, int y, int j, double2 v) const { apply(fld, k, y, j, v, load(fld, k, y, j)); }
    LG_D void finish_row(int fld, int k, int y) const {
        if (!pad) return;
        const long o = lay.at(k, y, nx);
        if (mode == 1 && k <= kmax[fld]) {
            // whole-row semantics of the Fortran array expressions on the two pad reals
            double2 vb = ld2(divt[fld] + o), vu = ld2(u[fld] + o);
            const double f = force[fld];
            double2 nr = make_double2(dadd(dsub(-0.0, vb.x), f), dadd(dsub(-0.0, vb.y), f));
            double2 vf;
            if (first_step) { vf = nr; *reinterpret_cast<double2*>(rhs_f[fld] + o) = nr; }
            else vf = ld2(rhs_f[fld] + o);
            *reinterpret_cast<double2*>(dst[fld] + o) = nr;
            *reinterpret_cast<double2*>(u[fld] + o) =
                make_double2(dadd(vu.x, dmul(dt, dadd(dmul(t1, nr.x), dmul(t2, vf.x)))),
                             dadd(vu.y, dmul(dt, dadd(dmul(t1, nr.y), dmul(t2, vf.y)))));
            return;
        }
        *reinterpret_cast<double2*>(dst[fld] + o) = make_double2(0.0, 0.0);
        if (mode == 2) {
            double2 vr = ld2(rhs_f[fld] + o);
            *reinterpret_cast<double2*>(rhs_f[fld] + o) = make_double2(dsub(vr.x, 0.0), dsub(vr.y, 0.0));
        }
    }
};

// ---- pointwise z kernels ------------------------------------------------------------
// dfdz(k) = (f(k+shift_hi) - f(k+shift_lo))/dz over 1:nx   (derivatives.f90:245-251, :292-298)
static __global__ void k_ddz(const double* __restrict__ f, double* __restrict__ dfdz, Lay lay, int nx, int ny,
                      int k0, int k1, int lo, int hi, double inv_dz) {
    const long n = long(nx / 2) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int j = int(t % (nx / 2));
        long r = t / (nx / 2);
        int y = int(r % ny), k = k0 + int(r / ny);
        double2 a = ld2(f + lay.at(k + hi, y, 2 * j)), b = ld2(f + lay.at(k + lo, y, 2 * j));
        *reinterpret_cast<double2*>(dfdz + lay.at(k, y, 2 * j)) =
            make_double2(dmul(inv_dz, dsub(a.x, b.x)), dmul(inv_dz, dsub(a.y, b.y)));
    }
}

// DNS wall derivatives, wallstress.f90:131-168: dudz(kdst) = sign*(u(ksrc) - uwall)/h with
// h = dz/2 (sign = +1 bottom, -1 top), dvdz(kdst) = sign*v(ksrc)/h, over 1:nx
static __global__ void k_wall_dns(const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ dudz,
                           double* __restrict__ dvdz, Lay lay, int nx, int ny, int ksrc, int kdst, double uwall,
                           double sign, double h) {
    const long n = long(nx / 2) * ny;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int j = int(t % (nx / 2)), y = int(t / (nx / 2));
        double2 a = ld2(u + lay.at(ksrc, y, 2 * j)), b = ld2(v + lay.at(ksrc, y, 2 * j));
        double2 du, dv;
        if (sign > 0) {
            du = make_double2(ddiv(dsub(a.x, uwall), h), ddiv(dsub(a.y, uwall), h));
            dv = make_double2(ddiv(b.x, h), ddiv(b.y, h));
        } else {
            du = make_double2(ddiv(dsub(uwall, a.x), h), ddiv(dsub(uwall, a.y), h));
            dv = make_double2(ddiv(-b.x, h), ddiv(-b.y, h));
        }
        *reinterpret_cast<double2*>(dudz + lay.at(kdst, y, 2 * j)) = du;
        *reinterpret_cast<double2*>(dvdz + lay.at(kdst, y, 2 * j)) = dv;
    }
}

// dst(k) = value on whole planes k0..k1-1 (BOGUS poisoning / zeroing)
static __global__ void k_fill(double* __restrict__ dst, long plane, int k0, int k1, double value) {
    const long n = plane * (k1 - k0);
    double* p = dst + long(k0) * plane;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) p[t] = value;
}

// several plane fills in one launch (blockIdx.y = entry): the BOGUS poisoning at the end of a routine is
// 3-6 single-plane fills, each a launch of its own otherwise (launch-bound on a many-rank slab)
struct FillList {
    static constexpr int kMax = 8;
    double* p[kMax];
    long n[kMax];
    double v[kMax];
    int count;
};
static __global__ void k_fill_multi(const FillList l) {
    const int e = blockIdx.y;
    double* p = l.p[e];
    const long n = l.n[e];
    const double v = l.v[e];
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) p[t] = v;
}

// max |f| over 1:nx, 1:ny, planes k0..k1-1 -> atomicMax on the bit pattern (values >= 0)
static __global__ void k_absmax(const double* __restrict__ f, Lay lay, int nx, int ny, int k0, int k1,
                         unsigned long long* __restrict__ out) {
    __shared__ double red[kBlock];
    const long n = long(nx / 2) * ny * (k1 - k0);
    double m = 0.0;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int j = int(t % (nx / 2));
        long r = t / (nx / 2);
        int y = int(r % ny), k = k0 + int(r / ny);
        double2 v = ld2(f + lay.at(k, y, 2 * j));
        m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(red[0]));
}

// sum |a+b+c| over 1:nx, 1:ny, planes k0..k1-1 (rmsdiv.f90:42-50); partial sums per block
static __global__ void k_abs3sum(const double* __restrict__ a, const double* __restrict__ b,
                          const double* __restrict__ c, Lay lay, int nx, int ny, int k0, int k1,
                          double* __restrict__ out) {
    __shared__ double red[kBlock];
    const long n = long(nx / 2) * ny * (k1 - k0);
    double s = 0.0;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int j = int(t % (nx / 2));
        long r = t / (nx / 2);
        int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, 2 * j);
        double2 va = ld2(a + o), vb = ld2(b + o), vc = ld2(c + o);
        s += fabs(va.x + vb.x + vc.x) + fabs(va.y + vb.y + vc.y);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if (int(threadIdx.x) < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

// ---- tridiagonal solve (tridag_array.f90 + press_stag_array.f90:149-239) ------------
// One thread per (kx, ky) mode marches the whole z column (n = nzt + 1 rows, row j <->
// p(:,:,j-1)).  The matrix is time-invariant: a = c = 1/dz^2, b = -(k2 + 2/dz^2), with
// the Neumann rows (b,c) = (-1, 1) at j = 1 and (a,b) = (-1, 1) at j = n, so gam(j) is
// tabulated once (k_tridag_setup) and bet(j) is re-derived in registers.  Division and
// multiply-subtract are kept un-contracted to follow tridag_array.f90:98-113 exactly.
struct TriGeom {
    int lh, ny, nzt;        // nzt = number of w levels = rows - 1
    int row;                // doubles per y row of the spectral arrays (ld)
    long plane;             // doubles per plane of the spectral arrays
    long gplane;            // doubles per plane of gam (lh * ny)
    double kxs, kys, dz;
};

LG_D double tri_k2(const TriGeom& g, int jx, int jy) {
    double kx = g.kxs * double(jx);
    double ky = g.kys * double(jy < g.ny / 2 ? jy : jy - g.ny);
    if (jy == g.ny / 2) ky = 0.0, kx = 0.0;
    return dadd(dmul(kx, kx), dmul(ky, ky));
}

// gam(j), j = 2..n  stored at gam[j*gplane + jy*lh + jx].  The pivots bet(j) depend only on (kx, ky, dz), so
// the reference's zero-pivot stop (tridag_array.f90:56-59,101-108, SAFETYMODE) is checked HERE, once, for the
// fused and pencil sweeps that reuse the table: *fail != 0 makes press_stag_array return an error.
static __global__ void k_tridag_setup(TriGeom g, double* __restrict__ gam, int* __restrict__ fail) {
    const int nm = (g.lh - 1) * g.ny;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nm) return;
    const int jx = t % (g.lh - 1), jy = t / (g.lh - 1);
    if (jy == g.ny / 2 || (jx == 0 && jy == 0)) return;
    const double c3 = ddiv(1.0, dmul(g.dz, g.dz));
    const double bb = -dadd(tri_k2(g, jx, jy), dmul(2.0, c3));
    const int n = g.nzt + 1;
    double bet = -1.0, cprev = 1.0;                 // row 1: b = -1, c = 1
    for (int j = 2; j <= n; ++j) {
        const double a = (j == n) ? -1.0 : c3;
        const double b = (j == n) ? 1.0 : bb;
        const double gm = ddiv(cprev, bet);
        bet = dsub(b, dmul(a, gm));
        if (bet == 0.0) *fail = 1;
        gam[long(j) * g.gplane + long(jy) * g.lh + jx] = gm;
        cprev = c3;
    }
}

// Forward + backward sweep with the right-hand side assembled on the fly from the
// spectra of H = u*/(tadv1 dt) (press_stag_array.f90:188-215); p_hat(k) = row k+1.
//   Hx, Hy, Hz : (ld, ny, 0:nzt) spectra; planes 1..nzt-1 of Hx,Hy and 1..nzt of Hz used
//   rbot, rtop : spectra of divtz at the walls (:114-126), one plane each
static __global__ void k_tridag_fused(TriGeom g, const double* __restrict__ Hx, const double* __restrict__ Hy,
                               const double* __restrict__ Hz, const double* __restrict__ rbot,
                               const double* __restrict__ rtop, const double* __restrict__ gam,
                               double* __restrict__ p) {
    const int nm = (g.lh - 1) * g.ny;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nm) return;
    const int jx = t % (g.lh - 1), jy = t / (g.lh - 1);
    const long mo = long(jy) * g.row + 2 * jx;      // offset of this mode inside a plane
    const int n = g.nzt + 1;
    const double dz = g.dz;
    if (jy == g.ny / 2) return;                      // never solved; zeroed by the consumer
    if (jx == 0 && jy == 0) {
        // zero-wavenumber chain, press_stag_array.f90:226-234
        double2 pk = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(p + mo) = pk;
        double2 rb = ld2(rbot + mo);
        pk = make_double2(dsub(pk.x, dmul(dz, rb.x)), dsub(pk.y, dmul(dz, rb.y)));
        *reinterpret_cast<double2*>(p + g.plane + mo) = pk;
        for (int k = 2; k <= g.nzt; ++k) {
            double2 h = ld2(Hz + long(k) * g.plane + mo);
            pk = make_double2(dadd(pk.x, dmul(h.x, dz)), dadd(pk.y, dmul(h.y, dz)));
            *reinterpret_cast<double2*>(p + long(k) * g.plane + mo) = pk;
        }
        return;
    }
    const double c3 = ddiv(1.0, dmul(dz, dz));
    const double c4 = ddiv(1.0, dz);
    const double kx = g.kxs * double(jx);
    const double ky = g.kys * double(jy < g.ny / 2 ? jy : jy - g.ny);
    const double bb = -dadd(dadd(dmul(kx, kx), dmul(ky, ky)), dmul(2.0, c3));
    // row 1: u(1) = r(1)/b(1) = (-dz*rbottomw)/(-1)
    double2 rb = ld2(rbot + mo);
    double2 u = make_double2(ddiv(dmul(-dz, rb.x), -1.0), ddiv(dmul(-dz, rb.y), -1.0));
    *reinterpret_cast<double2*>(p + mo) = u;
    double bet = -1.0;
    double2 hzm = ld2(Hz + g.plane + mo);            // Hz(j-1) for j = 2
    for (int j = 2; j <= n; ++j) {
        double a, b;
        double2 r;
        if (j < n) {
            a = c3; b = bb;
            double2 hx = ld2(Hx + long(j - 1) * g.plane + mo);
            double2 hy = ld2(Hy + long(j - 1) * g.plane + mo);
            double2 hz = ld2(Hz + long(j) * g.plane + mo);
            // aH_x + aH_y + (rH_z(j) - rH_z(j-1))*const4
            r.x = dadd(dadd(dmul(-hx.y, kx), dmul(-hy.y, ky)), dmul(dsub(hz.x, hzm.x), c4));
            r.y = dadd(dadd(dmul(hx.x, kx), dmul(hy.x, ky)), dmul(dsub(hz.y, hzm.y), c4));
            hzm = hz;
        } else {
            a = -1.0; b = 1.0;
            double2 rt = ld2(rtop + mo);
            r = make_double2(dmul(-dz, rt.x), dmul(-dz, rt.y));
        }
        const double gm = gam[long(j) * g.gplane + long(jy) * g.lh + jx];
        bet = dsub(b, dmul(a, gm));
        u = make_double2(ddiv(dsub(r.x, dmul(a, u.x)), bet), ddiv(dsub(r.y, dmul(a, u.y)), bet));
        *reinterpret_cast<double2*>(p + long(j - 1) * g.plane + mo) = u;
    }
    // back substitution, tridag_array.f90:141-152 (j = n-1 .. 1)
    for (int j = n - 1; j >= 1; --j) {
        const double gm = gam[long(j + 1) * g.gplane + long(jy) * g.lh + jx];
        double2 uj = ld2(p + long(j - 1) * g.plane + mo);
        u = make_double2(dsub(uj.x, dmul(gm, u.x)), dsub(uj.y, dmul(gm, u.y)));
        *reinterpret_cast<double2*>(p + long(j - 1) * g.plane + mo) = u;
    }
}

// dpdz(k) = (p(k) - p(k-1))/dz over 1:nx (press_stag_array.f90:284-288): true division.
// With rhs != nullptr also RHSz -= dpdz (main.f90:323) and, for kproj <= k < kproj_end,
// w += dt*(-tadv1*dpdz) (forcing.f90:195-207).
static __global__ void k_dpdz(const double* __restrict__ p, double* __restrict__ dpdz, Lay lay, int nx, int ny,
                              int k0, int k1, double dz, double* __restrict__ rhs, double* __restrict__ w,
                              int kproj, int kproj_end, double dt, double t1) {
    const long n = long(nx / 2) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int j = int(t % (nx / 2));
        long r = t / (nx / 2);
        int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, 2 * j);
        double2 a = ld2(p + o), b = ld2(p + lay.at(k - 1, y, 2 * j));
        double2 v = make_double2(ddiv(dsub(a.x, b.x), dz), ddiv(dsub(a.y, b.y), dz));
        *reinterpret_cast<double2*>(dpdz + o) = v;
        if (rhs) {
            double2 vr = ld2(rhs + o);
            *reinterpret_cast<double2*>(rhs + o) = make_double2(dsub(vr.x, v.x), dsub(vr.y, v.y));
            if (k >= kproj && k < kproj_end) {
                double2 vu = ld2(w + o);
                *reinterpret_cast<double2*>(w + o) = make_double2(dadd(vu.x, dmul(dt, dmul(-t1, v.x))),
                                                                  dadd(vu.y, dmul(dt, dmul(-t1, v.y))));
            }
        }
    }
}

// padd (fft.f90:43-71) / unpadd (fft.f90:74-99) as plain spectral-array copies, for callers
// that use them directly (scalars.f90); the convective term fuses them into the y pass.
static __global__ void k_padd(const double* __restrict__ u, double* __restrict__ ub, int nx, int ny, int ld,
                              int ny2, int ld_big, int nplanes) {
    const int lhb = ld_big / 2;
    const long n = long(lhb) * ny2 * nplanes;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int m = int(t % lhb);
        long r = t / lhb;
        int jb = int(r % ny2), k = int(r / ny2);
        int js = -1;
        if (jb < ny / 2) js = jb;
        else if (jb > ny2 - ny / 2) js = jb - (ny2 - ny);
        double2 v = make_double2(0.0, 0.0);
        if (js >= 0 && 2 * m < nx) v = ld2(u + (long(k) * ny + js) * ld + 2 * m);
        *reinterpret_cast<double2*>(ub + (long(k) * ny2 + jb) * ld_big + 2 * m) = v;
    }
}
static __global__ void k_unpadd(double* __restrict__ cc, const double* __restrict__ cb, int nx, int ny, int ld,
                                int ny2, int ld_big, int nplanes) {
    const int lh = ld / 2;
    const long n = long(lh) * ny * nplanes;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        int m = int(t % lh);
        long r = t / lh;
        int j = int(r % ny), k = int(r / ny);
        double2 v = make_double2(0.0, 0.0);
        if (j != ny / 2 && 2 * m < nx) {
            int jb = j < ny / 2 ? j : j + (ny2 - ny);
            v = ld2(cb + (long(k) * ny2 + jb) * ld_big + 2 * m);
        }
        *reinterpret_cast<double2*>(cc + (long(k) * ny + j) * ld + 2 * m) = v;
    }
}

// tridag_array with caller-supplied coefficients (tridag_array.f90:166-246, serial form):
// a,b,c (lh, ny, n), r,u (ld, ny, n); gam is an (lh, ny, n) work array.
static __global__ void k_tridag_general(int lh, int ny, int n, int ld, const double* __restrict__ a,
                                        const double* __restrict__ b, const double* __restrict__ c,
                                        const double* __restrict__ r, double* __restrict__ u,
                                        double* __restrict__ gam, int* __restrict__ fail) {
    const int nm = (lh - 1) * ny;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nm) return;
    const int jx = t % (lh - 1), jy = t / (lh - 1);
    const long cp = long(lh) * ny, rp = long(ld) * ny;
    const long co = long(jy) * lh + jx, ro = long(jy) * ld + 2 * jx;
    // row 1 for every jy, jx <= lh-1 (:193-205)
    double bet = b[co];
    if (bet == 0.0) { *fail = 1; return; }
    double2 r1 = ld2(r + ro);
    double2 uu = make_double2(ddiv(r1.x, bet), ddiv(r1.y, bet));
    *reinterpret_cast<double2*>(u + ro) = uu;
    if (jy == ny / 2 || (jx == 0 && jy == 0)) return;
    for (int j = 1; j < n; ++j) {
        const double gm = ddiv(c[(j - 1) * cp + co], bet);
        gam[j * cp + co] = gm;
        const double aj = a[j * cp + co];
        bet = dsub(b[j * cp + co], dmul(aj, gm));
        if (bet == 0.0) { *fail = 1; return; }
        double2 rj = ld2(r + j * rp + ro);
        uu = make_double2(ddiv(dsub(rj.x, dmul(aj, uu.x)), bet), ddiv(dsub(rj.y, dmul(aj, uu.y)), bet));
        *reinterpret_cast<double2*>(u + j * rp + ro) = uu;
    }
    for (int j = n - 2; j >= 0; --j) {
        const double gm = gam[(j + 1) * cp + co];
        double2 uj = ld2(u + j * rp + ro);
        uu = make_double2(dsub(uj.x, dmul(gm, uu.x)), dsub(uj.y, dmul(gm, uu.y)));
        *reinterpret_cast<double2*>(u + j * rp + ro) = uu;
    }
}

// ---- multi-rank pressure solve: z-slabs -> (kx,ky)-pencils and back -----------------------------
// The tridiagonal systems couple all z-slabs.  Each rank assembles the right-hand-side rows it
// owns, an all-to-all turns slabs into pencils (every rank gets ALL rows of ny/nproc ky-rows,
// the layout mpi_transpose_mod.f90 sketches), the Thomas sweep runs un-split with exactly the
// arithmetic of tridag_array.f90, and a second all-to-all returns p_hat.  Row ownership: rank 0
// owns local rows 1..nz, every other rank rows 2..nz+1 (row nz+1 is real only on the top rank),
// i.e. `nz` block rows per rank; block row i of rank r is local row i + (r ? 2 : 1).
struct PencilGeom {
    int lh, ny, ld, nz, nproc, coord;
    int cy;                 // ky rows per pencil chunk = ceil(ny / nproc); the last rank's chunk may be ragged or empty
    long plane;             // ld * ny
    double kxs, kys, dz;
    // peer-memory transposes (lesgo_gpu_comm_p2p_import): pencil[q] is rank q's pencil buffer of this solve,
    // mapped into this rank's address space (NVLink P2P); p2p = 0: NCCL path
    double* pencil[8];
    int p2p;
    LG_HD long block() const { return long(nz) * cy * ld; }     // doubles per (src, dst) block
};

// send layout: buf[dst q][block row i][jy_local][ld]
static __global__ void k_press_pack(PencilGeom g, const double* __restrict__ Hx, const double* __restrict__ Hy,
                                    const double* __restrict__ Hz, const double* __restrict__ rbot,
                                    const double* __restrict__ rtop, double* __restrict__ buf) {
    const int lhm = g.lh - 1;
    const long n = long(lhm) * g.ny * g.nz;
    const double c4 = ddiv(1.0, g.dz);
    const bool bottom = g.coord == 0, top = g.coord == g.nproc - 1;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int jx = int(t % lhm);
        long r = t / lhm;
        const int jy = int(r % g.ny), i = int(r / g.ny);
        const int j = i + (bottom ? 1 : 2);                       // local system row
        const long mo = long(jy) * g.ld + 2 * jx;
        double2 v;
        if (j == 1) {                                             // bottom Neumann row (:160)
            double2 rb = ld2(rbot + mo);
            v = make_double2(dmul(-g.dz, rb.x), dmul(-g.dz, rb.y));
        } else if (j == g.nz + 1) {                               // top Neumann row (:173), top rank only
            if (top) { double2 rt = ld2(rtop + mo); v = make_double2(dmul(-g.dz, rt.x), dmul(-g.dz, rt.y)); }
            else v = make_double2(0.0, 0.0);
        } else if (jx == 0 && jy == 0) {
            v = ld2(Hz + long(j) * g.plane + mo);                 // k=0 chain needs H_z itself (:232)
        } else {
            const double kx = g.kxs * double(jx);
            const double ky = g.kys * double(jy < g.ny / 2 ? jy : jy - g.ny);
            double2 hx = ld2(Hx + long(j - 1) * g.plane + mo), hy = ld2(Hy + long(j - 1) * g.plane + mo);
            double2 hz = ld2(Hz + long(j) * g.plane + mo), hzm = ld2(Hz + long(j - 1) * g.plane + mo);
            v.x = dadd(dadd(dmul(-hx.y, kx), dmul(-hy.y, ky)), dmul(dsub(hz.x, hzm.x), c4));
            v.y = dadd(dadd(dmul(hx.x, kx), dmul(hy.x, ky)), dmul(dsub(hz.y, hzm.y), c4));
        }
        const int q = jy / g.cy, jl = jy % g.cy;
        // NCCL path: into the local send buffer, block q.  Peer-memory path: the assembly kernel IS the
        // transpose -- the value goes straight into block `coord` of rank q's pencil buffer over NVLink.
        double* dst = g.p2p ? g.pencil[q] + g.coord * g.block() : buf + q * g.block();
        *reinterpret_cast<double2*>(dst + (long(i) * g.cy + jl) * g.ld + 2 * jx) = v;
    }
}

// Stream-ordered barrier of the peer-memory transposes without a collective: every rank's kernel stores the epoch
// into slot [rank] of EVERY rank's signal array (remote stores over NVLink, after a system-scope fence: the preceding
// kernel of this stream -- the one whose remote stores / in-place sweep the barrier protects -- has completed) and
// then waits until all nproc slots of its OWN array carry that epoch.  One block of 32 threads; replaces a one-double
// ncclAllReduce (~50 us at 8 GPUs) per barrier.
struct P2PSig {
    unsigned long long* peer[8];
    int rank, nproc;
    unsigned long long epoch;
};
static __global__ void k_p2p_barrier(P2PSig sg) {
    const int q = threadIdx.x;
#ifdef LESGO_EMUL
    if (q == 0) {      // the emulator runs a block's threads one after the other: signal everybody first, then wait
        for (int p = 0; p < sg.nproc; ++p) __atomic_store_n(&sg.peer[p][sg.rank], sg.epoch, __ATOMIC_RELEASE);
        for (int p = 0; p < sg.nproc; ++p)
            while (__atomic_load_n(&sg.peer[sg.rank][p], __ATOMIC_ACQUIRE) < sg.epoch) sched_yield();
    }
#else
    __threadfence_system();
    if (q < sg.nproc) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sg.peer[q] + sg.rank), "l"(sg.epoch) : "memory");
        unsigned long long v = 0;
        const unsigned long long* mine = sg.peer[sg.rank] + q;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        } while (v < sg.epoch);
    }
    __syncwarp();
    __threadfence_system();
#endif
}

// gam table in pencil layout: gam[global row][jy_local][jx]
static __global__ void k_tridag_setup_pencil(PencilGeom g, int nzt, double* __restrict__ gam, int* __restrict__ fail) {
    const int nm = (g.lh - 1) * g.cy;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nm) return;
    const int jx = t % (g.lh - 1), jl = t / (g.lh - 1), jy = g.coord * g.cy + jl;
    if (jy >= g.ny || jy == g.ny / 2 || (jx == 0 && jy == 0)) return;
    const double c3 = ddiv(1.0, dmul(g.dz, g.dz));
    const double kx = g.kxs * double(jx);
    const double ky = g.kys * double(jy < g.ny / 2 ? jy : jy - g.ny);
    const double bb = -dadd(dadd(dmul(kx, kx), dmul(ky, ky)), dmul(2.0, c3));
    const int n = nzt + 1;
    double bet = -1.0, cprev = 1.0;
    for (int j = 2; j <= n; ++j) {
        const double a = (j == n) ? -1.0 : c3;
        const double b = (j == n) ? 1.0 : bb;
        const double gm = ddiv(cprev, bet);
        bet = dsub(b, dmul(a, gm));
        if (bet == 0.0) *fail = 1;
        gam[(long(j) * g.cy + jl) * g.lh + jx] = gm;
        cprev = c3;
    }
}

// recv layout: buf[src r][block row i][jy_local][ld]; solved in place.  The matrix is real, so the
// real and imaginary parts of a mode are independent systems: one thread each (same arithmetic,
// twice the parallelism for the short pencils of a many-rank run).
template <bool PIPE>
static __global__ void k_tridag_pencil(PencilGeom g, int nzt, const double* __restrict__ gam, double* __restrict__ buf) {
    const int nm = (g.lh - 1) * g.cy;
    const int t2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (t2 >= 2 * nm) return;
    const int t = t2 >> 1, part = t2 & 1;
    const int jx = t % (g.lh - 1), jl = t / (g.lh - 1), jy = g.coord * g.cy + jl;
    if (jy >= g.ny || jy == g.ny / 2) return;
    const int n = nzt + 1;
    const long mo = long(jl) * g.ld + 2 * jx + part;
    const long rs = long(g.cy) * g.ld;                            // doubles between block rows
    // global row gj (1..n) -> address
    auto at = [&](int gj) -> double* {
        int r, i;
        if (gj <= g.nz) { r = 0; i = gj - 1; }
        else {
            r = (gj - 2) / (g.nz - 1);
            if (r > g.nproc - 1) r = g.nproc - 1;
            i = gj - r * (g.nz - 1) - 2;
        }
        return buf + r * g.block() + i * rs + mo;
    };
    if (jx == 0 && jy == 0) {
        // zero-wavenumber chain (press_stag_array.f90:226-234): row 1 holds -dz*rbottomw = p(1),
        // rows 2..nzt hold H_z(0,0,k); p(k) is system row k+1.
        double carry = *at(1);                                    // p(1) = 0 - dz*rbottomw
        *at(1) = 0.0;                                             // p(0) = 0
        for (int k = 2; k <= nzt; ++k) {
            const double h = *at(k);
            *at(k) = carry;                                       // row k = p(k-1)
            carry = dadd(carry, dmul(h, g.dz));
        }
        *at(n) = carry;                                           // row n = p(nzt)
        return;
    }
    const double c3 = ddiv(1.0, dmul(g.dz, g.dz));
    const double kx = g.kxs * double(jx);
    const double ky = g.kys * double(jy < g.ny / 2 ? jy : jy - g.ny);
    const double bb = -dadd(dadd(dmul(kx, kx), dmul(ky, ky)), dmul(2.0, c3));
    double u = ddiv(*at(1), -1.0);
    *at(1) = u;
    double bet = -1.0;
    // The recurrence is serial in j but its operands are not: fetch UN rows ahead so the sweep pays one
    // memory latency per UN rows instead of one per row (a many-rank run has too few modes per GPU to
    // hide it with other threads).  PIPE: the operands of batch i+1 are requested BEFORE batch i is
    // processed (two register sets), so that latency also overlaps the division chain of the batch in hand.
    // Same arithmetic, same order.
    constexpr int UN = 8;
    double* pj[2][UN];
    double r[2][UN], gm[2][UN];
    auto ldf = [&](const int s, int j0) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            const int j = j0 + q;
            if (j <= n) { pj[s][q] = at(j); r[s][q] = *pj[s][q]; gm[s][q] = gam[(long(j) * g.cy + jl) * g.lh + jx]; }
        }
    };
    auto cpf = [&](const int s, int j0) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            const int j = j0 + q;
            if (j <= n) {
                const double a = (j == n) ? -1.0 : c3;
                const double b = (j == n) ? 1.0 : bb;
                bet = dsub(b, dmul(a, gm[s][q]));
                u = ddiv(dsub(r[s][q], dmul(a, u)), bet);
                *pj[s][q] = u;
            }
        }
    };
    if (PIPE) {
        ldf(0, 2);
        for (int j0 = 2; j0 <= n; j0 += 2 * UN) {
            ldf(1, j0 + UN);
            cpf(0, j0);
            ldf(0, j0 + 2 * UN);
            cpf(1, j0 + UN);
        }
    } else {
        for (int j0 = 2; j0 <= n; j0 += UN) { ldf(0, j0); cpf(0, j0); }
    }
    auto ldb = [&](const int s, int j0) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            const int j = j0 - q;
            if (j >= 1) { pj[s][q] = at(j); r[s][q] = *pj[s][q]; gm[s][q] = gam[(long(j + 1) * g.cy + jl) * g.lh + jx]; }
        }
    };
    auto cpb = [&](const int s, int j0) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            const int j = j0 - q;
            if (j >= 1) {
                u = dsub(r[s][q], dmul(gm[s][q], u));
                *pj[s][q] = u;
            }
        }
    };
    if (PIPE) {
        // the first backward batch re-reads rows the forward sweep has just written: same thread, program order
        ldb(0, n - 1);
        for (int j0 = n - 1; j0 >= 1; j0 -= 2 * UN) {
            ldb(1, j0 - UN);
            cpb(0, j0);
            ldb(0, j0 - 2 * UN);
            cpb(1, j0 - UN);
        }
    } else {
        for (int j0 = n - 1; j0 >= 1; j0 -= UN) { ldb(0, j0); cpb(0, j0); }
    }
}

// after the return all-to-all: buf[src q = ky chunk][block row i][jy_local][ld] -> p_hat planes
static __global__ void k_press_unpack(PencilGeom g, const double* __restrict__ buf, double* __restrict__ p) {
    const long n = long(g.lh) * g.ny * g.nz;
    const bool bottom = g.coord == 0;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int jx = int(t % g.lh);
        long r = t / g.lh;
        const int jy = int(r % g.ny), i = int(r / g.ny);
        const int k = i + (bottom ? 0 : 1);                       // plane = local row - 1
        const int q = jy / g.cy, jl = jy % g.cy;
        // NCCL path: block q of the local return buffer.  Peer-memory path: PULL block `coord` of rank q's
        // pencil buffer, where its Thomas sweep left p_hat in place (this kernel has the parallelism to hide
        // the NVLink load latency; the latency-bound sweep keeps purely local accesses)
        const double* src = g.p2p ? g.pencil[q] + g.coord * g.block() : buf + q * g.block();
        double2 v = ld2(src + (long(i) * g.cy + jl) * g.ld + 2 * jx);
        *reinterpret_cast<double2*>(p + long(k) * g.plane + long(jy) * g.ld + 2 * jx) = v;
    }
}

// Fused time-stepping glue (one pass instead of two per component):
//  F_RHS_AB2      : rhs = -rhs - divt + force                         main.f90:211-214,229-232
//                   [first step: rhs_f = rhs                           main.f90:273-280]
//                   u = u + dt*(tadv1*rhs + tadv2*rhs_f)               main.f90:287-296
//  F_GRADP_PROJECT: rhs = rhs - dpd                                    main.f90:321-326
//                   u = u + dt*(-tadv1*dpd)  for k >= kproj, over 1:nx forcing.f90:171-207
enum FusedMode { F_RHS_AB2 = 0, F_GRADP_PROJECT = 1 };
static __global__ void k_glue_fused(int mode, double* __restrict__ rhs, const double* __restrict__ b,
                                    double* __restrict__ rhs_f, double* __restrict__ u, Lay lay, int nx, int ny,
                                    int k0, int k1, int kproj, int first_step, double force, double dt, double t1,
                                    double t2) {
    const int half = lay.row / 2;
    const long n = long(half) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int j = int(t % half);
        long r = t / half;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, 2 * j);
        double2 vr = ld2(rhs + o), vb = ld2(b + o), vu = ld2(u + o);
        if (mode == F_RHS_AB2) {
            double2 nr = make_double2(dadd(dsub(-vr.x, vb.x), force), dadd(dsub(-vr.y, vb.y), force));
            double2 vf;
            if (first_step) { vf = nr; *reinterpret_cast<double2*>(rhs_f + o) = nr; }
            else vf = ld2(rhs_f + o);
            *reinterpret_cast<double2*>(rhs + o) = nr;
            *reinterpret_cast<double2*>(u + o) =
                make_double2(dadd(vu.x, dmul(dt, dadd(dmul(t1, nr.x), dmul(t2, vf.x)))),
                             dadd(vu.y, dmul(dt, dadd(dmul(t1, nr.y), dmul(t2, vf.y)))));
        } else {
            *reinterpret_cast<double2*>(rhs + o) = make_double2(dsub(vr.x, vb.x), dsub(vr.y, vb.y));
            if (k >= kproj && 2 * j < nx)
                *reinterpret_cast<double2*>(u + o) = make_double2(dadd(vu.x, dmul(dt, dmul(-t1, vb.x))),
                                                                  dadd(vu.y, dmul(dt, dmul(-t1, vb.y))));
        }
    }
}

// ---- wall stress, strain rate, constant-coefficient SGS stress, stress divergence ---
struct SgsParams {
    int nz, bottom, top, lbc_mom, ubc_mom, sgs;
    double nu;          // nu_molec/(u_star z_i) when molec, else 0   (sgs_param.f90:188-192)
    double Cs_opt2;     // Co**2 (sgs_model 1) or 0.03 (dynamic models before DYN_init)
};

// calc_Sij (sgs_stag_util.f90:467-634) + |S| and Nu_t (:226-235) on planes k0..k1-1 (1 <= k <= nz).
// S[6] = S11, S12, S13, S22, S23, S33; lsq[k] = l(k)**2.
struct SijArgs {
    const double *dudx, *dudy, *dudz, *dvdx, *dvdy, *dvdz, *dwdx, *dwdy, *dwdz;
    double* S[6];
    double* Nu_t;
    const double* lsq;
};
static __global__ void k_sij_nut(SijArgs a, SgsParams p, Lay lay, int nx, int ny, int k0, int k1) {
    const long n = long(nx) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, i), om = lay.at(k - 1, y, i);
        double s11, s12, s13, s22, s23, s33;
        if (p.bottom && k == 1) {
            if (p.lbc_mom == 0) {
                s11 = a.dudx[o]; s12 = dmul(0.5, dadd(a.dudy[o], a.dvdx[o])); s13 = dmul(0.5, dadd(a.dudz[o], a.dwdx[o]));
                s22 = a.dvdy[o]; s23 = dmul(0.5, dadd(a.dvdz[o], a.dwdy[o])); s33 = dmul(0.5, dadd(a.dwdz[o], 0.0));
            } else {
                const long o2 = lay.at(2, y, i);
                s11 = a.dudx[o]; s12 = dmul(0.5, dadd(a.dudy[o], a.dvdx[o]));
                const double wx = dmul(0.5, dadd(a.dwdx[o], a.dwdx[o2]));
                s13 = dmul(0.5, dadd(a.dudz[o], wx));
                s22 = a.dvdy[o];
                const double wy = dmul(0.5, dadd(a.dwdy[o], a.dwdy[o2]));
                s23 = dmul(0.5, dadd(a.dvdz[o], wy));
                s33 = a.dwdz[o];
            }
        } else if (p.top && k == p.nz) {
            if (p.ubc_mom == 0) {
                s11 = a.dudx[om]; s12 = dmul(0.5, dadd(a.dudy[om], a.dvdx[om])); s13 = dmul(0.5, dadd(a.dudz[o], a.dwdx[o]));
                s22 = a.dvdy[om]; s23 = dmul(0.5, dadd(a.dvdz[o], a.dwdy[o])); s33 = dmul(0.5, dadd(a.dwdz[om], 0.0));
            } else {
                s11 = a.dudx[om]; s12 = dmul(0.5, dadd(a.dudy[om], a.dvdx[om]));
                const double wx = dmul(0.5, dadd(a.dwdx[om], a.dwdx[o]));
                s13 = dmul(0.5, dadd(a.dudz[o], wx));
                s22 = a.dvdy[om];
                const double wy = dmul(0.5, dadd(a.dwdy[om], a.dwdy[o]));
                s23 = dmul(0.5, dadd(a.dvdz[o], wy));
                s33 = a.dwdz[om];
            }
        } else {
            s11 = dmul(0.5, dadd(a.dudx[o], a.dudx[om]));
            const double uy = dadd(a.dudy[o], a.dudy[om]), vx = dadd(a.dvdx[o], a.dvdx[om]);
            s12 = dmul(0.25, dadd(uy, vx));
            s13 = dmul(0.5, dadd(a.dudz[o], a.dwdx[o]));
            s22 = dmul(0.5, dadd(a.dvdy[o], a.dvdy[om]));
            s23 = dmul(0.5, dadd(a.dvdz[o], a.dwdy[o]));
            s33 = dmul(0.5, dadd(a.dwdz[o], a.dwdz[om]));
        }
        a.S[0][o] = s11; a.S[1][o] = s12; a.S[2][o] = s13; a.S[3][o] = s22; a.S[4][o] = s23; a.S[5][o] = s33;
        double nut = 0.0;
        if (p.sgs) {
            const double q = dadd(dadd(dadd(dmul(s11, s11), dmul(s22, s22)), dmul(s33, s33)),
                                  dmul(2.0, dadd(dadd(dmul(s12, s12), dmul(s13, s13)), dmul(s23, s23))));
            nut = dmul(dmul(sqrt(dmul(2.0, q)), p.Cs_opt2), a.lsq[k]);
        }
        a.Nu_t[o] = nut;
    }
}

// tau_ij (sgs_stag_util.f90:237-428) on planes k0..k1-1 (1 <= k <= nz-1); T[6] = txx, txy, txz, tyy, tyz, tzz
struct TauArgs {
    const double* S[6];
    const double* Nu_t;
    double* T[6];
};
static __global__ void k_tau(TauArgs a, SgsParams p, Lay lay, int nx, int ny, int k0, int k1) {
    const long n = long(nx) * ny * (k1 - k0);
    // S index of the four "uvp-node" stresses txx, txy, tyy, tzz and their T index
    const int sI[4] = {0, 1, 3, 5}, tI[4] = {0, 1, 3, 5};
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx);
        long r = t / nx;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, i), op = lay.at(k + 1, y, i);
        const double nu = p.nu;
        if (p.bottom && k == 1) {
            // txz, tyz of plane 1 come from wallstress
            if (p.lbc_mom == 0) {
                const double cst = p.sgs ? dadd(dmul(0.5, dadd(a.Nu_t[o], a.Nu_t[op])), nu) : nu;
                for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(-cst, dadd(a.S[sI[q]][o], a.S[sI[q]][op]));
            } else {
                const double cst = p.sgs ? dmul(-2.0, dadd(a.Nu_t[o], nu)) : dmul(-2.0, nu);
                for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(cst, a.S[sI[q]][o]);
            }
        } else if (p.top && k == p.nz - 1) {
            if (p.ubc_mom == 0) {
                const double cst = p.sgs ? dadd(dmul(0.5, dadd(a.Nu_t[o], a.Nu_t[op])), nu) : nu;
                const double cst2 = p.sgs ? dmul(2.0, dadd(a.Nu_t[o], nu)) : dmul(2.0, nu);
                for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(-cst, dadd(a.S[sI[q]][o], a.S[sI[q]][op]));
                a.T[2][o] = dmul(-cst2, a.S[2][o]);
                a.T[4][o] = dmul(-cst2, a.S[4][o]);
            } else if (p.sgs) {
                const double cst = dmul(-2.0, dadd(a.Nu_t[op], nu)), cst2 = dmul(-2.0, dadd(a.Nu_t[o], nu));
                for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(cst, a.S[sI[q]][op]);
                a.T[2][o] = dmul(cst2, a.S[2][o]);
                a.T[4][o] = dmul(cst2, a.S[4][o]);
            } else {
                // the reference's DNS branch uses Sij(nz-1) here (:353-356)
                for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(dmul(-2.0, nu), a.S[sI[q]][o]);
                a.T[2][o] = dmul(dmul(-2.0, nu), a.S[2][o]);
                a.T[4][o] = dmul(dmul(-2.0, nu), a.S[4][o]);
            }
        } else if (p.sgs) {
            const double c3 = dmul(dmul(-2.0, nu), 0.5), c4 = dmul(-2.0, nu);
            const double cst = dmul(-0.5, dadd(a.Nu_t[o], a.Nu_t[op])), cst2 = dmul(-2.0, a.Nu_t[o]);
            for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(dadd(cst, c3), dadd(a.S[sI[q]][o], a.S[sI[q]][op]));
            a.T[2][o] = dmul(dadd(cst2, c4), a.S[2][o]);
            a.T[4][o] = dmul(dadd(cst2, c4), a.S[4][o]);
        } else {
            for (int q = 0; q < 4; ++q) a.T[tI[q]][o] = dmul(-nu, dadd(a.S[sI[q]][o], a.S[sI[q]][op]));
            a.T[2][o] = dmul(dmul(-2.0, nu), a.S[2][o]);
            a.T[4][o] = dmul(dmul(-2.0, nu), a.S[4][o]);
        }
    }
}

// wallstress.f90:131-168 stresses of the DNS walls: t = -nu * d(u|v)/dz on plane k (1:nx)
static __global__ void k_wall_tau_dns(const double* __restrict__ dudz, const double* __restrict__ dvdz,
                                      double* __restrict__ txz, double* __restrict__ tyz, Lay lay, int nx, int ny,
                                      int k, double nu) {
    const long n = long(nx) * ny;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const long o = lay.at(k, int(t / nx), int(t % nx));
        txz[o] = dmul(-nu, dudz[o]);
        tyz[o] = dmul(-nu, dvdz[o]);
    }
}

// equilibrium wall model, wallstress.f90:171-255.  u1, v1: test-filtered u, v of the wall-adjacent
// plane (one plane each); sign = +1 bottom (ksrc = 1, kdst = 1), -1 top (ksrc = nz-1, kdst = nz).
static __global__ void k_wall_equil(const double* __restrict__ u, const double* __restrict__ v,
                                    const double* __restrict__ u1, const double* __restrict__ v1,
                                    double* __restrict__ dudz, double* __restrict__ dvdz, double* __restrict__ txz,
                                    double* __restrict__ tyz, Lay lay, int nx, int ny, int ksrc, int kdst, double sign,
                                    double vonk, double denom, double hk /* 0.5*dz*vonk */) {
    const long n = long(nx) * ny;
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int i = int(t % nx), y = int(t / nx);
        const long os = lay.at(ksrc, y, i), od = lay.at(kdst, y, i), o1 = long(y) * lay.row + i;
        const double a = u1[o1], b = v1[o1];
        const double uavg = sqrt(dadd(dmul(a, a), dmul(b, b)));
        const double ustar = ddiv(dmul(uavg, vonk), denom);
        const double cst = dmul(-sign, ddiv(dmul(ustar, ustar), uavg));
        txz[od] = dmul(cst, a);
        tyz[od] = dmul(cst, b);
        const double uu = u[os], vv = v[os];
        const double g = dmul(sign, ddiv(ustar, hk));
        dudz[od] = uu == 0.0 ? 0.0 : ddiv(dmul(g, uu), uavg);
        dvdz[od] = vv == 0.0 ? 0.0 : ddiv(dmul(g, vv), uavg);
    }
}

// out = (a + b) [+ c] over 1:nx of planes k0..k1-1, pad columns zeroed when zero_pad
// (divstress_uv.f90:66-83, divstress_w.f90:74-114)
static __global__ void k_sum3(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                              const double* __restrict__ c, Lay lay, int nx, int ny, int k0, int k1, int zero_pad) {
    const int half = lay.row / 2;
    const long n = long(half) * ny * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const int j = int(t % half);
        long r = t / half;
        const int y = int(r % ny), k = k0 + int(r / ny);
        const long o = lay.at(k, y, 2 * j);
        if (2 * j >= nx) {
            if (zero_pad) *reinterpret_cast<double2*>(out + o) = make_double2(0.0, 0.0);
            continue;
        }
        double2 va = ld2(a + o), vb = ld2(b + o);
        double2 s2 = make_double2(dadd(va.x, vb.x), dadd(va.y, vb.y));
        if (c) { double2 vc = ld2(c + o); s2 = make_double2(dadd(s2.x, vc.x), dadd(s2.y, vc.y)); }
        *reinterpret_cast<double2*>(out + o) = s2;
    }
}

}  // namespace lg
