// prodfwd.cu -- launcher of convec's products + forward x transform on the 3/2 grid (prodfwd_kernels.h)
#include "launch.h"
#include "prodfwd_kernels.h"
#include "sizes.h"
namespace lg {
template <int NX2>
static int launch_prodfwd_n(const ProdArgs& a, const cplx* W, const cplx* Wh, cudaStream_t s) {
    typedef ProdCfg<NX2> C;
    LG_SET_SMEM((k_prodfwd<NX2>), C::smem);
    const long nwork = long(a.ny2) * a.nchunks;
    if (nwork <= 0) return 0;
    dim3 grid(persistent_blocks(C::smem, nwork, C::MINB));
    LG_LAUNCH((k_prodfwd<NX2>), grid, dim3(C::NTHR), C::smem, s, a, W, Wh);
    return 0;
}
#define LG_PRODFWD_CASE(S, B) case B: return launch_prodfwd_n<B>(a, W, Wh, s);
int launch_prodfwd(int nx2, const ProdArgs& a, const cplx* W, const cplx* Wh, cudaStream_t s) {
    switch (nx2) { LG_SIZE_PAIRS(LG_PRODFWD_CASE) }
    return -1;
}
}  // namespace lg
