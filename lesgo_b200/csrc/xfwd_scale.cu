#include "xfwd_impl.h"
namespace lg {
#define PRO ProScale
template <>
int launch_xfwd<ProScale>(int NX, const ProScale& pro, int nfields, const XfOut& out, int ny, int k0,
                          int nplanes, const cplx* W, const cplx* Wh, cudaStream_t s) {
    switch (NX) {
        LG_SIZE_PAIRS(LG_XFWD_CASE_SMALL)
        case 24: return launch_xfwd_n<24, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 72: return launch_xfwd_n<72, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 120: return launch_xfwd_n<120, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 144: return launch_xfwd_n<144, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 240: return launch_xfwd_n<240, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 288: return launch_xfwd_n<288, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 480: return launch_xfwd_n<480, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 576: return launch_xfwd_n<576, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 768: return launch_xfwd_n<768, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
        case 1536: return launch_xfwd_n<1536, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
    }
    return -1;
}
}  // namespace lg
