#include "launch.h"
#include "sizes.h"
namespace lg {
template <int NX, class Epi = EpiStore>
static int launch_xinv_n(const XiSrc& in, const Epi& epi, int nfields, int ny, int k0, int nplanes,
                         const cplx* W, const cplx* Wh, cudaStream_t s, bool big = false) {
    typedef XCfg<NX> C;
    const long nrows = long(ny) * nplanes;
    if (nrows <= 0) return 0;
    if (warp_passes() == 3) {
        // warp-scope with cp.async prefetch of each warp's next row (rows of >= 256 complex points)
        typedef XWCfg<NX, false, true> CW;
        if (CW::PREF) {
            LG_SET_SMEM((k_xinv_w<NX, Epi, false, true>), CW::smem);
            const long nwork = nrows * nfields;
            dim3 grid(persistent_blocks(CW::smem, (nwork + CW::WPB - 1) / CW::WPB, CW::MINB));
            LG_LAUNCH((k_xinv_w<NX, Epi, false, true>), grid, dim3(CW::NTHR), CW::smem, s, in, epi, nfields, ny, k0, nplanes, W, Wh);
            return 0;
        }
    }
    if (warp_passes() == 2 || ((warp_passes() == 1 || warp_passes() == 3) && big)) {
        typedef XWCfg<NX> CW;
        LG_SET_SMEM((k_xinv_w<NX, Epi>), CW::smem);
        const long nwork = ((nrows + CW::NF - 1) / CW::NF) * nfields;
        dim3 grid(persistent_blocks(CW::smem, (nwork + CW::WPB - 1) / CW::WPB, CW::MINB));
        LG_LAUNCH((k_xinv_w<NX, Epi>), grid, dim3(CW::NTHR), CW::smem, s, in, epi, nfields, ny, k0, nplanes, W, Wh);
        return 0;
    }
    LG_SET_SMEM((k_xinv<NX, Epi>), C::smem);
    dim3 grid(persistent_blocks(C::smem, ((nrows + C::NF - 1) / C::NF) * nfields, C::MINB));
    LG_LAUNCH((k_xinv<NX, Epi>), grid, dim3(C::NTHR), C::smem, s, in, epi, nfields, 0, ny, k0, nplanes, W, Wh);
    return 0;
}
#define LG_XINV_CASE_SMALL(S, B) case S: return launch_xinv_n<S>(in, epi, nfields, ny, k0, nplanes, W, Wh, s);
#define LG_XINV_BIG(B) case B: return launch_xinv_n<B>(in, epi, nfields, ny, k0, nplanes, W, Wh, s, true);
int launch_xinv(int NX, const XiSrc& in, const EpiStore& epi, int nfields, int ny, int k0, int nplanes,
                const cplx* W, const cplx* Wh, cudaStream_t s) {
    switch (NX) {
        LG_SIZE_PAIRS(LG_XINV_CASE_SMALL)
        LG_XINV_BIG(24)
        LG_XINV_BIG(72)
        LG_XINV_BIG(120)
        LG_XINV_BIG(144)
        LG_XINV_BIG(240)
        LG_XINV_BIG(288)
        LG_XINV_BIG(480)
        LG_XINV_BIG(576)
        LG_XINV_BIG(768)
        LG_XINV_BIG(1536)
    }
    return -1;
}
#define LG_XINVF_CASE(S, B) case S: return launch_xinv_n<S, EpiFused>(in, epi, nfields, ny, k0, nplanes, W, Wh, s);
int launch_xinv_fused(int NX, const XiSrc& in, const EpiFused& epi, int nfields, int ny, int k0, int nplanes,
                      const cplx* W, const cplx* Wh, cudaStream_t s) {
    switch (NX) { LG_SIZE_PAIRS(LG_XINVF_CASE) }
    return -1;
}
}  // namespace lg
