// shared body of the xfwd_*.cu instantiation files
#pragma once
#include "launch.h"
#include "sizes.h"

namespace lg {
template <class Pro> constexpr bool zmajor_for() { return false; }
template <> constexpr bool zmajor_for<ProConvec>() { return true; }
template <int NX, class Pro>
static int launch_xfwd_n(const Pro& pro, int nfields, const XfOut& out, int ny, int k0, int nplanes,
                         const cplx* W, const cplx* Wh, cudaStream_t s) {
    typedef XCfg<NX> C;
    const long nrows = long(ny) * nplanes;
    if (nrows <= 0) return 0;
    if (warp_passes() == 2) {
        typedef XWCfg<NX> CW;
        LG_SET_SMEM((k_xfwd_w<NX, Pro>), CW::smem);
        const long nwork = ((nrows + CW::NF - 1) / CW::NF) * nfields;
        dim3 grid(persistent_blocks(CW::smem, (nwork + CW::WPB - 1) / CW::WPB, CW::MINB));
        LG_LAUNCH((k_xfwd_w<NX, Pro>), grid, dim3(CW::NTHR), CW::smem, s, pro, out, nfields, ny, k0, nplanes, W, Wh);
        return 0;
    }
    LG_SET_SMEM((k_xfwd<NX, Pro>), C::smem);
    dim3 grid(persistent_blocks(C::smem, ((nrows + C::NF - 1) / C::NF) * nfields, XFMinB<Pro, C>::value));
    LG_LAUNCH((k_xfwd<NX, Pro>), grid, dim3(C::NTHR), C::smem, s, pro, out, nfields, int(zmajor_for<Pro>() && ny % C::NF == 0), ny, k0, nplanes, W, Wh);
    return 0;
}
}  // namespace lg

#define LG_XFWD_CASE_SMALL(S, B) case S: return launch_xfwd_n<S, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
#define LG_XFWD_CASE_BIG(S, B) case B: return launch_xfwd_n<B, PRO>(pro, nfields, out, ny, k0, nplanes, W, Wh, s);
