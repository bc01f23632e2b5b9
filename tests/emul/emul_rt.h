// emul_rt.h -- CPU emulation of the small CUDA subset the lesgo_b200 kernels use.
// TEST INFRASTRUCTURE ONLY (see lesgo_b200/csrc/portable.h).  One OS thread runs one
// block at a time; the block's CUDA threads are ucontext fibres scheduled round-robin
// between barriers, which reproduces __syncthreads() semantics exactly.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <sched.h>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

namespace emu {
struct ThreadCtx { uint3 tid, bid; dim3 bdim, gdim; unsigned char* smem; };
extern thread_local ThreadCtx* cur;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void syncthreads();
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)
#define __syncthreads() emu::syncthreads()
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define LG_HD inline __attribute__((always_inline))
#define LG_D inline __attribute__((always_inline))
#define __ldg(p) (*(p))
#define LG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define LG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::cur->smem)

// ---- the sliver of the CUDA runtime API the host code calls -------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < h; ++i) std::memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8

// atomics used by the reduction kernels (blocks may run on several OS threads)
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline double atomicAdd(double* a, double v) {
    double old = *a, neu;
    do { neu = old + v; } while (!__atomic_compare_exchange(a, &old, &neu, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
