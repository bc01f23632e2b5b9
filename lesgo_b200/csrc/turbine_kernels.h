// turbine_kernels.h -- actuator-disk forcing: turbines.f90:465-638 on the device.
// The host hands over, once, the node lists and indicator weights turbines_nodes builds
// (turbines.f90:275-462); per step the disks gather their velocity, the scalar update of every disk
// runs in one small block, and the forces are scattered back -- no host round trip.
#pragma once
#include "ops.h"

namespace lg {

struct TurbSet {
    int nloc;
    const int* start;        // (nloc + 1) prefix offsets into the node arrays
    const long* off;         // per node: offset of (i, j, k) in a (ld, ny, 0:nz) field
    const double* ind;       // per node: indicator weight
    const int* owner;        // per node: 1 when this entry is the LAST one written at its grid point
                             // (the reference assigns, so a later disk overwrites an earlier one)
    const double* nhat;      // (nloc, 3)
    const double* Ct_prime;  // (nloc)
    const double* dia;
    const double* M;
    double* u_d;             // (nloc) out: disk-averaged velocity
    double* u_d_T;           // (nloc) inout: its running average
    double* f_n;             // (nloc) out
    // ADM with rotation (use_rotation, turbines.f90:607-615): per node tangential weight and unit vector, or null
    const double* ind_t;     // per node
    const double* e_theta;   // per node, 3 components
    double tip_speed_ratio;
};

// turbines.f90:521-548: disk_avg_vel(s) = sum_l dx dy dz ind(l) (nhat . (u, v, w_uv)); one block per disk
static __global__ void k_turb_gather(TurbSet t, const double* __restrict__ u, const double* __restrict__ v,
                                     const double* __restrict__ w, long plane, double vol) {
    __shared__ double red[kBlock];
    const int s = blockIdx.x;
    const double n0 = t.nhat[3 * s], n1 = t.nhat[3 * s + 1], n2 = t.nhat[3 * s + 2];
    double acc = 0.0;
    for (int l = t.start[s] + threadIdx.x; l < t.start[s + 1]; l += blockDim.x) {
        const long o = t.off[l];
        const double w_uv = dmul(0.5, dadd(w[o + plane], w[o]));               // functions.f90:77
        const double un = dadd(dadd(dmul(n0, u[o]), dmul(n1, v[o])), dmul(n2, w_uv));
        acc = dadd(acc, dmul(dmul(vol, t.ind[l]), un));
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int h = kBlock / 2; h > 0; h >>= 1) {
        if (int(threadIdx.x) < h) red[threadIdx.x] = dadd(red[threadIdx.x], red[threadIdx.x + h]);
        __syncthreads();
    }
    if (threadIdx.x == 0) t.u_d[s] = red[0];
}

// turbines.f90:570-588: ADM correction, first-order time filter, thrust per unit mass
static __global__ void k_turb_update(TurbSet t, double eps, int adm_correction) {
    for (int s = threadIdx.x; s < t.nloc; s += blockDim.x) {
        double ud = t.u_d[s];
        const double Ct = t.Ct_prime[s];
        if (adm_correction) ud = ddiv(ud, dadd(1.0, dmul(dmul(0.25, dsub(1.0, t.M[s])), Ct)));
        const double udT = dadd(dmul(dsub(1.0, eps), t.u_d_T[s]), dmul(eps, ud));
        const double d = t.dia[s];
        t.u_d[s] = ud;
        t.u_d_T[s] = udT;
        const double q = dmul(dmul(dmul(dmul(dmul(-0.5, Ct), fabs(udT)), udT), 0.25), 3.14159265358979323846);
        t.f_n[s] = dmul(q, dmul(d, d));
    }
}

// turbines.f90:599-615: f = f_n nhat ind (+ f_n e_theta ind_t / tip_speed_ratio with use_rotation) at the disk's nodes
// (fz still on uv nodes)
static __global__ void k_turb_scatter(TurbSet t, double* __restrict__ fxa, double* __restrict__ fya, double* __restrict__ fz_uv) {
    const int s = blockIdx.x;
    const double fn = t.f_n[s];
    const double f0 = dmul(fn, t.nhat[3 * s]), f1 = dmul(fn, t.nhat[3 * s + 1]), f2 = dmul(fn, t.nhat[3 * s + 2]);
    for (int l = t.start[s] + threadIdx.x; l < t.start[s + 1]; l += blockDim.x) {
        if (!t.owner[l]) continue;
        const long o = t.off[l];
        const double a = t.ind[l];
        double gx = dmul(f0, a), gy = dmul(f1, a), gz = dmul(f2, a);
        if (t.ind_t) {
            const double b = t.ind_t[l];
            gx = dadd(gx, ddiv(dmul(dmul(fn, t.e_theta[3 * l]), b), t.tip_speed_ratio));
            gy = dadd(gy, ddiv(dmul(dmul(fn, t.e_theta[3 * l + 1]), b), t.tip_speed_ratio));
            gz = dadd(gz, ddiv(dmul(dmul(fn, t.e_theta[3 * l + 2]), b), t.tip_speed_ratio));
        }
        fxa[o] = gx; fya[o] = gy; fz_uv[o] = gz;
    }
}

// interp_to_w_grid (functions.f90:97-141): out(k) = (in(k-1) + in(k)) / 2 on planes k0..k1-1, whole rows
static __global__ void k_interp_w(const double* __restrict__ in, double* __restrict__ out, long plane, int k0, int k1) {
    const long n = plane * (k1 - k0);
    for (long t = long(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += long(gridDim.x) * blockDim.x) {
        const long o = long(k0) * plane + t;
        out[o] = dmul(0.5, dadd(in[o - plane], in[o]));
    }
}

}  // namespace lg
