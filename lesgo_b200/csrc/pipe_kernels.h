// pipe_kernels.h -- plane pipeline: the x-forward, y and x-inverse passes of ONE routine in ONE
// persistent kernel, so the spectral intermediates between the passes never leave L2.
//
// Why: with one launch per pass over the whole slab every 2-D transform writes its x-pass
// intermediate to HBM and reads it back (measured round 2: the core step moves 107 GB against
// 19.9 GB algorithmic, and with the FFT arithmetic removed the passes still take 21 of 29 ms --
// the step is bound by that traffic, profiles/r2_nocompute.md).  Here the work items of all
// passes are dealt from one global ticket counter in DEPENDENCY ORDER, plane by plane:
//
//     step s:  [x-inverse tiles of plane s-2] [y tiles of plane s-1] [x-forward tiles of plane s]
//
// A block takes the next ticket, waits (acquire) until the planes it consumes are complete,
// runs the ordinary tile routine (xfwd_work / ypass_work / xinv_work of fft_kernels.h) and
// publishes (release) a per-plane completion count.  The intermediates live in RINGS of `ring`
// planes (slot = plane % ring) that are overwritten while still resident in the 126 MB L2, so
// the only HBM traffic left is the routine's inputs and outputs.  A producer re-using a ring
// slot waits for the consumer of the plane that occupied it.
//
// Deadlock freedom: every dependency of a ticket has a smaller ticket number and tickets are
// handed out in order, so the block holding the oldest unfinished ticket never waits; blocks
// that are not resident hold no ticket.
#pragma once
#include "ops.h"

namespace lg {

struct PipeCtl {
    unsigned* ticket;      // zeroed before the launch
    int* done;             // 3 * nplanes completion counters (x-forward, y, x-inverse), zeroed
    int nplanes, k0;       // planes k0 .. k0+nplanes-1
    int ring;              // ring depth of the intermediates, in planes
    int nf_f, nf_i;        // fields of the x-forward / x-inverse phases
    int ny_f, ny_i;        // rows per plane of the x passes
    int i_k0[kMaxFields], i_k1[kMaxFields];   // x-inverse field f is produced on planes [i_k0, i_k1) only
};

#ifdef LESGO_EMUL
LG_HD int ld_acquire(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
LG_HD void red_release(int* p) { __atomic_fetch_add(p, 1, __ATOMIC_RELEASE); }
LG_HD void spin_pause() { sched_yield(); }
#else
LG_D int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
LG_D void red_release(int* p) { asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory"); }
LG_D void spin_pause() { __nanosleep(100); }
#endif

// all threads of the block: wait until *ctr >= target
LG_D void pipe_wait(const int* ctr, int target) {
    if (threadIdx.x == 0) {
        while (ld_acquire(ctr) < target) spin_pause();
    }
    __syncthreads();
}
// all threads of the block: the tile's global writes are complete -> count it
LG_D void pipe_signal(int* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) red_release(ctr);
}

struct NoPro {
    LG_D double2 load(int, int, int, int) const { return make_double2(0.0, 0.0); }
};
struct NoEpi {
    LG_D void store(int, int, int, int, double2) const {}
    LG_D void finish_row(int, int, int) const {}
};

template <int NXF, int NIN, int NOUT, bool MULTI, int NXI>
struct PipeCfg {
    static constexpr bool HF = NXF > 0, HI = NXI > 0;
    static_assert(!(HF && HI) || NXF == NXI, "both x passes of a pipeline have the same length");
    static constexpr int NX = HF ? NXF : NXI;
    typedef XCfg<NX> CX;
    typedef YCfg<NIN, NOUT, MULTI> CY;
    static constexpr int NTHR = CY::NTHR;
    static constexpr bool ok = (CX::NTHR == CY::NTHR);    // the tile routines assume blockDim == their NTHR
    static constexpr int XBUF = CX::NF * CX::SL, YBUF = CY::NBUF * CY::TC * CY::SL;
    static constexpr int BUF = XBUF > YBUF ? XBUF : YBUF;
    static constexpr int XTW = CX::TWL + CX::NWH, YTW = CY::TWI + CY::TWO;
    static constexpr size_t smem = size_t(BUF + XTW + YTW) * sizeof(cplx);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int by_regs = 65536 / (NTHR * 96) < 1 ? 1 : 65536 / (NTHR * 96);
    static constexpr int MINB = by_smem < by_regs ? by_smem : by_regs;
};

template <int NXF, class Pro, int NIN, int NOUT, bool MULTI, int NXI, class Epi>
__global__ void __launch_bounds__(PipeCfg<NXF, NIN, NOUT, MULTI, NXI>::NTHR, PipeCfg<NXF, NIN, NOUT, MULTI, NXI>::MINB)
k_pipe(const __grid_constant__ Pro pro, const __grid_constant__ XfOut xo, const __grid_constant__ YArgs ya,
       const __grid_constant__ XiSrc xi, const __grid_constant__ Epi epi, const __grid_constant__ PipeCtl ctl,
       const cplx* __restrict__ Wxg, const cplx* __restrict__ Whxg, const cplx* __restrict__ Wing,
       const cplx* __restrict__ Woutg) {
    typedef PipeCfg<NXF, NIN, NOUT, MULTI, NXI> C;
    typedef typename C::CX CX;
    typedef typename C::CY CY;
    constexpr bool HF = C::HF, HI = C::HI;
    LG_DYN_SMEM(cplx, sm);
    cplx* buf = sm;
    cplx* S = MULTI ? sm + CY::TC * CY::SL : sm;
    cplx* Wx = sm + C::BUF;
    cplx* Whx = Wx + CX::TWL;
    cplx* Win = Whx + CX::NWH;
    cplx* Wout = Win + CY::TWI;
    __shared__ int s_k[CX::NF], s_y[CX::NF];
    __shared__ double* s_p[CX::NF];
    __shared__ unsigned s_t;
    load_table(Wx, Wxg, CX::TWL);
    load_table(Whx, Whxg, CX::NWH);
    if (NIN > 0) load_table(Win, Wing, CY::TWI);
    if (NOUT > 0) load_table(Wout, Woutg, CY::TWO);

    const int np = ctl.nplanes;
    const int tpf = HF ? ((ctl.ny_f + CX::NF - 1) / CX::NF) * ctl.nf_f : 0;     // tiles per plane
    const int tpy = ((ya.ncols + CY::TC - 1) / CY::TC) * ya.nfields;
    const int tpi = HI ? ((ctl.ny_i + CX::NF - 1) / CX::NF) * ctl.nf_i : 0;
    const int lag_y = HF ? 1 : 0, lag_i = lag_y + 1;
    const unsigned nsteps = unsigned(np + (HI ? lag_i : lag_y));
    const unsigned tps = unsigned(tpf + tpy + tpi);
    int* done_f = ctl.done;
    int* done_y = done_f + np;
    int* done_i = done_y + np;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = atomicAdd(ctl.ticket, 1u);
        __syncthreads();
        const unsigned t = s_t;
        const unsigned s = t / tps;
        int r = int(t % tps);
        if (s >= nsteps) break;
        if (r < tpi) {
            if constexpr (HI) {
                const int p = int(s) - lag_i;
                if (p < 0 || p >= np) continue;
                pipe_wait(done_y + p, tpy);
                const int fld = r % ctl.nf_i, k = ctl.k0 + p;
                if (k >= ctl.i_k0[fld] && k < ctl.i_k1[fld])
                    xinv_work<(HI ? NXI : 16), Epi>(buf, Wx, Whx, s_k, s_y, s_p, xi, epi, ctl.nf_i, ctl.ny_i, ctl.k0, np,
                                                   unsigned(r) + unsigned(p) * unsigned(tpi));
                pipe_signal(done_i + p);
            }
        } else if (r < tpi + tpy) {
            r -= tpi;
            const int p = int(s) - lag_y;
            if (p < 0 || p >= np) continue;
            if (HF) pipe_wait(done_f + p, tpf);
            if (HI && p >= ctl.ring) pipe_wait(done_i + (p - ctl.ring), tpi);      // ring slot of the outputs is free
            ypass_work<NIN, NOUT, MULTI>(buf, S, Win, Wout, ya, unsigned(r) + unsigned(p) * unsigned(tpy));
            pipe_signal(done_y + p);
        } else {
            if constexpr (HF) {
                r -= tpi + tpy;
                const int p = int(s);
                if (p >= np) continue;
                if (p >= ctl.ring) pipe_wait(done_y + (p - ctl.ring), tpy);        // ring slot of the x spectra is free
                xfwd_work<(HF ? NXF : 16), Pro>(buf, Wx, Whx, s_k, s_y, s_p, pro, xo, ctl.nf_f, ctl.ny_f, ctl.k0, np,
                                               unsigned(r) + unsigned(p) * unsigned(tpf));
                pipe_signal(done_f + p);
            }
        }
    }
}

}  // namespace lg
