#include <cstdlib>
#include "launch.h"
#include "wfft2_kernels.h"
#include "sizes.h"
namespace lg {
// LESGO_XW2 bit mask: 1 = the 3/2-grid x inverse (measured on B200 at 512 x 512 x 256: 2.81 -> 2.11 ms), 2 = small-grid
// rows with the plain epilogue, 4 = small-grid rows with the fused time-stepping epilogues.  Default 1: on the small
// grid the two-stage kernel's 12 warps per SM hide the epilogues' operand loads worse than the 24 of the three-stage
// one (6.69 against 5.35 ms for the six small-grid launches), profiles/r4_experiments.md.
int xw2_mask() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_XW2"); v = e ? std::atoi(e) : 1; }
    return v;
}
bool xw2_enabled() { return xw2_mask() != 0; }
template <class Epi> struct IsFusedEpi { static constexpr bool value = false; };
template <> struct IsFusedEpi<EpiFused> { static constexpr bool value = true; };
bool xplan2_lookup(int m, int* r1, int* r2) {
#define LG_XP2(M_) if (m == M_ && XPlan2<M_>::on) { *r1 = XPlan2<M_>::R1; *r2 = XPlan2<M_>::R2; return true; }
    LG_XP2(256) LG_XP2(384)
#undef LG_XP2
    return false;
}
template <int NX, class Epi = EpiStore>
static int launch_xinv_n(const XiSrc& in, const Epi& epi, int nfields, int ny, int k0, int nplanes,
                         const cplx* W, const cplx* Wh, cudaStream_t s, bool big = false) {
    typedef XCfg<NX> C;
    const long nrows = long(ny) * nplanes;
    if (nrows <= 0) return 0;
    if constexpr (XPlan2<NX / 2>::on) {
        // two-stage warp-scope transform, a half-warp per row (wfft2_kernels.h); LESGO_XW2=0 keeps the three-stage one
        if (warp_passes() == 3 && (xw2_mask() & (big ? 1 : (IsFusedEpi<Epi>::value ? 4 : 2)))) {
            typedef XW2Cfg<NX> C2;
            LG_SET_SMEM((k_xinv_w2<NX, Epi>), C2::smem);
            const long nwork = ((nrows + 1) / 2) * nfields;
            dim3 grid(persistent_blocks(C2::smem, (nwork + C2::WPB - 1) / C2::WPB, C2::MINB));
            LG_LAUNCH((k_xinv_w2<NX, Epi>), grid, dim3(C2::NTHR), C2::smem, s, in, epi, nfields, ny, k0, nplanes,
                      W + PlanInfo<NX / 2>::twlen, Wh);
            return 0;
        }
    }
    if (warp_passes() == 3) {
        // warp-scope with cp.async prefetch of each warp's next row (rows of >= 256 complex points)
        typedef XWCfg<NX, false, true> CW;
        if (CW::PREF) {
            LG_SET_SMEM((k_xinv_w<NX, Epi, false, true>), CW::smem);
            const long nwork = nrows * nfields;
            dim3 grid(persistent_blocks(CW::smem, (nwork + CW::WPB - 1) / CW::WPB, XWMinB<Epi, CW>::value));
            LG_LAUNCH((k_xinv_w<NX, Epi, false, true>), grid, dim3(CW::NTHR), CW::smem, s, in, epi, nfields, ny, k0, nplanes, W, Wh);
            return 0;
        }
    }
    if (warp_passes() == 2 || ((warp_passes() == 1 || warp_passes() == 3) && big)) {
        typedef XWCfg<NX> CW;
        LG_SET_SMEM((k_xinv_w<NX, Epi>), CW::smem);
        const long nwork = ((nrows + CW::NF - 1) / CW::NF) * nfields;
        dim3 grid(persistent_blocks(CW::smem, (nwork + CW::WPB - 1) / CW::WPB, XWMinB<Epi, CW>::value));
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
