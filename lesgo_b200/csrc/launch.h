// launch.h -- size-dispatching launchers (explicitly instantiated in x*.cu / ypass.cu so
// the many template instantiations compile in parallel).
#pragma once
#include <atomic>
#include "ops.h"
#include "wfft_kernels.h"

namespace lg {

// returns 0, or -1 if the length is not in sizes.h
template <class Pro>
int launch_xfwd(int NX, const Pro& pro, int nfields, const XfOut& out, int ny, int k0, int nplanes,
                const cplx* W, const cplx* Wh, cudaStream_t s);
int launch_xinv(int NX, const XiSrc& in, const EpiStore& epi, int nfields, int ny, int k0, int nplanes,
                const cplx* W, const cplx* Wh, cudaStream_t s);
// small-grid lengths only (the fused time-stepping epilogues of lesgo_gpu_step)
int launch_xinv_fused(int NX, const XiSrc& in, const EpiFused& epi, int nfields, int ny, int k0, int nplanes,
                      const cplx* W, const cplx* Wh, cudaStream_t s);
int launch_ypass(int nin, int nout, const YArgs& a, int nfields, int nplanes, const cplx* Win,
                 const cplx* Wout, cudaStream_t s);
// derivative pass of length ny that also emits the padded (3 ny / 2) inverse transform into fld[].out2
int launch_ypass_pad2(int ny, const YArgs& a, int nfields, int nplanes, const cplx* Ws, const cplx* Wb, cudaStream_t s);
bool size_supported(int n_small);
// radices and stage-twiddle-table length of the plan for complex length n (false if none)
bool plan_lookup(int n, PlanDesc* out);
// two-stage column plan (fft_core.h Plan2) of the y passes for length n, if one exists
bool plan2_lookup(int n, int* r1, int* r2);
// two-stage warp-scope x-inverse plan (wfft2_kernels.h) for rows of m complex points, if one exists
bool xplan2_lookup(int m, int* r1, int* r2);
bool xw2_enabled();      // LESGO_XW2=0 switches the two-stage x inverse off

// products + forward x transform of convec, marching up z (prodfwd_kernels.h)
struct ProdArgs;
int launch_prodfwd(int nx2, const ProdArgs& a, const cplx* W, const cplx* Wh, cudaStream_t s);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several
// GPUs (ranks as threads) must set it once on each of them.  Every call site keeps its own
// bit mask of the device ordinals already served (one static per template instantiation).
#ifdef LESGO_EMUL
#define LG_SET_SMEM(kernel, bytes) do { (void)(bytes); } while (0)
#else
#define LG_SET_SMEM(kernel, bytes)                                                                  \
    do {                                                                                            \
        static std::atomic<unsigned long long> done_{0ull};                                         \
        int dev_ = 0;                                                                               \
        cudaGetDevice(&dev_);                                                                       \
        const unsigned long long bit_ = 1ull << (dev_ & 63);                                        \
        if (!(done_.load(std::memory_order_acquire) & bit_)) {                                      \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));  \
            done_.fetch_or(bit_, std::memory_order_release);                                        \
        }                                                                                           \
    } while (0)
#endif

// persistent grids: blocks per SM that fit (shared memory / 64-register budget), times SMs
int sm_count();
// warp-scope x passes (wfft_kernels.h), LESGO_XW: 0 never; 1 for the 3/2-grid x inverse only; 2 everywhere;
// 3 (default) every x inverse of >= 256 complex points at warp scope WITH cp.async prefetch of each warp's
// next row (5.34 + 2.79 ms against 5.66 + 2.87 ms), the 3/2-grid x inverse at warp scope otherwise
int warp_passes();
inline int persistent_blocks(size_t smem_bytes, long ntiles, int max_per_sm) {
    int per_sm = int((227 * 1024) / (smem_bytes + 1024));
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long g = long(sm_count()) * per_sm;
    return int(g < ntiles ? g : (ntiles < 1 ? 1 : ntiles));
}

}  // namespace lg
