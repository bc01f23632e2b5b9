// portable.h -- one source, two builds.
//
//  * Product build (nvcc, -gencode arch=compute_100a,code=sm_100a): the macros below
//    vanish and the kernels are ordinary CUDA.
//  * Kernel-logic emulator (g++ -DLESGO_EMUL, tests/emul/): the SAME kernel source is
//    compiled for the host and every CUDA thread of a block runs as a fibre so that
//    __syncthreads(), shared memory and threadIdx behave exactly as on the device.
//    That build exists only so index maps, wall-plane special cases and the C-ABI
//    host logic can be unit-tested in the CPU-only container; it is test
//    infrastructure (like oracle/), is never loaded by lesgo_b200 and is not a
//    fallback: the product library refuses to initialise without a CUDA device.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef LESGO_EMUL
#include "../../tests/emul/emul_rt.h"
#else
#include <cuda_runtime.h>
#define LG_HD __host__ __device__ __forceinline__
#define LG_D __device__ __forceinline__
#define LG_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// dynamic shared memory base (16-byte aligned)
#define LG_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char _lg_dyn_smem[]; \
    type* name = reinterpret_cast<type*>(_lg_dyn_smem)
#endif
