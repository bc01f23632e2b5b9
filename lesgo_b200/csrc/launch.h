// launch.h -- size-dispatching launchers (explicitly instantiated in x*.cu / ypass.cu so
// the many template instantiations compile in parallel).
#pragma once
#include "ops.h"

namespace lg {

// returns 0, or -1 if the length is not in sizes.h
template <class Pro>
int launch_xfwd(int NX, const Pro& pro, int nfields, const XfOut& out, int ny, int k0, int nplanes,
                const cplx* W, const cplx* Wh, cudaStream_t s);
int launch_xinv(int NX, const XiSrc& in, const EpiStore& epi, int nfields, int ny, int k0, int nplanes,
                const cplx* W, const cplx* Wh, cudaStream_t s);
int launch_ypass(int nin, int nout, const YArgs& a, int nfields, int nplanes, const cplx* Win,
                 const cplx* Wout, cudaStream_t s);
bool size_supported(int n_small);

template <class K> inline void set_smem(K kernel, size_t bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

}  // namespace lg
