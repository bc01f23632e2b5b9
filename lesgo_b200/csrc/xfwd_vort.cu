#include "xfwd_impl.h"
namespace lg {
#define PRO ProVort
template <>
int launch_xfwd<ProVort>(int NX, const ProVort& pro, int nfields, const XfOut& out, int ny, int k0,
                         int nplanes, const cplx* W, const cplx* Wh, cudaStream_t s) {
    switch (NX) { LG_SIZE_PAIRS(LG_XFWD_CASE_SMALL) }
    return -1;
}
}  // namespace lg
