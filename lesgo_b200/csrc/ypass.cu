#include <cstdlib>
#include "launch.h"
#include "sizes.h"
namespace lg {
template <int NIN, int NOUT, bool MULTI>
static int launch_y_m(const YArgs& a0, int nfields, int nplanes, const cplx* Win, const cplx* Wout,
                      cudaStream_t s) {
    typedef YCfg<NIN, NOUT, MULTI> C;
    LG_SET_SMEM((k_ypass<NIN, NOUT, MULTI>), C::smem);
    if (nplanes <= 0 || nfields <= 0) return 0;
    YArgs a = a0;
    a.nplanes = nplanes;
    a.nfields = nfields;
    const long ntiles = long((a.ncols + C::TC - 1) / C::TC) * nplanes;
    dim3 grid(persistent_blocks(C::smem, ntiles * nfields, C::MINB));
    LG_LAUNCH((k_ypass<NIN, NOUT, MULTI>), grid, dim3(C::NTHR), C::smem, s, a, Win, Wout);
    return 0;
}
template <int NIN, int NOUT>
static int launch_y_n(const YArgs& a, int nfields, int nplanes, const cplx* Win, const cplx* Wout,
                      cudaStream_t s) {
    if constexpr (NOUT > 0 && (NIN == NOUT || NIN == 0)) {
        // number of inverse transforms per field: an IKX output rides on the COPY transform
        int ntr = 0;
        bool has_copy = false, has_ikx = false;
        for (int o = 0; o < a.nout; ++o) {
            const int m = a.fld[0].out[o].mode;
            if (m == Y_COPY) has_copy = true; else if (m == Y_IKX) has_ikx = true; else ++ntr;
        }
        ntr += (has_copy || has_ikx) ? 1 : 0;
        if (ntr > 1) return launch_y_m<NIN, NOUT, true>(a, nfields, nplanes, Win, Wout, s);
    }
    return launch_y_m<NIN, NOUT, false>(a, nfields, nplanes, Win, Wout, s);
}
#define LG_Y_CASES(S, B)                                                                     \
    if (nin == S && nout == S) return launch_y_n<S, S>(a, nfields, nplanes, Win, Wout, s);   \
    if (nin == S && nout == 0) return launch_y_n<S, 0>(a, nfields, nplanes, Win, Wout, s);   \
    if (nin == 0 && nout == S) return launch_y_n<0, S>(a, nfields, nplanes, Win, Wout, s);   \
    if (nin == S && nout == B) return launch_y_n<S, B>(a, nfields, nplanes, Win, Wout, s);   \
    if (nin == B && nout == S) return launch_y_n<B, S>(a, nfields, nplanes, Win, Wout, s);
int launch_ypass(int nin, int nout, const YArgs& a, int nfields, int nplanes, const cplx* Win,
                 const cplx* Wout, cudaStream_t s) {
    LG_SIZE_PAIRS(LG_Y_CASES)
    // raw transforms on the 3/2 grid (forw_big / back_big of fft.f90:118-121)
#define LG_Y_RAW(B)                                                                          \
    if (nin == B && nout == 0) return launch_y_n<B, 0>(a, nfields, nplanes, Win, Wout, s);   \
    if (nin == 0 && nout == B) return launch_y_n<0, B>(a, nfields, nplanes, Win, Wout, s);
    LG_Y_RAW(24) LG_Y_RAW(72) LG_Y_RAW(120) LG_Y_RAW(144) LG_Y_RAW(240) LG_Y_RAW(288) LG_Y_RAW(480)
    LG_Y_RAW(576) LG_Y_RAW(768) LG_Y_RAW(1536)
    return -1;
}
// same-size multi-output pass that ALSO writes the 3/2-rule padded inverse transform (fld[].out2)
template <int NS_, int NB_>
static int launch_y_pad2(const YArgs& a0, int nfields, int nplanes, const cplx* Ws, const cplx* Wb, cudaStream_t s) {
    typedef YCfg<NS_, NS_, true, NB_> C;
    LG_SET_SMEM((k_ypass<NS_, NS_, true, NB_>), C::smem);
    if (nplanes <= 0 || nfields <= 0) return 0;
    YArgs a = a0;
    a.nplanes = nplanes;
    a.nfields = nfields;
    const long ntiles = long((a.ncols + C::TC - 1) / C::TC) * nplanes;
    dim3 grid(persistent_blocks(C::smem, ntiles * nfields, C::MINB));
    LG_LAUNCH((k_ypass<NS_, NS_, true, NB_>), grid, dim3(C::NTHR), C::smem, s, a, Ws, Ws, Wb);
    return 0;
}
#define LG_Y_PAD2(S, B) if (ny == S) return launch_y_pad2<S, B>(a, nfields, nplanes, Ws, Wb, s);
int launch_ypass_pad2(int ny, const YArgs& a, int nfields, int nplanes, const cplx* Ws, const cplx* Wb, cudaStream_t s) {
    LG_SIZE_PAIRS(LG_Y_PAD2)
    return -1;
}
#define LG_SUP(S, B) if (n == S) return true;
bool plan2_lookup(int n, int* r1, int* r2) {
#define LG_PL2(N) if (n == N && Plan2<N>::on) { *r1 = Plan2<N>::R1; *r2 = Plan2<N>::R2; return true; }
    LG_PL2(256) LG_PL2(384) LG_PL2(512) LG_PL2(768)
#undef LG_PL2
    return false;
}
bool plan_lookup(int n, PlanDesc* out) {
#define LG_PL(N) if (n == N) { *out = plan_desc<N>(); return true; }
    LG_PL(8) LG_PL(12) LG_PL(16) LG_PL(24) LG_PL(32) LG_PL(36) LG_PL(40) LG_PL(48) LG_PL(60) LG_PL(64) LG_PL(72)
    LG_PL(80) LG_PL(96) LG_PL(120) LG_PL(128) LG_PL(144) LG_PL(160) LG_PL(192) LG_PL(240) LG_PL(256) LG_PL(288)
    LG_PL(320) LG_PL(384) LG_PL(480) LG_PL(512) LG_PL(576) LG_PL(640) LG_PL(768) LG_PL(1024) LG_PL(1536)
#undef LG_PL
    return false;
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
#ifdef LESGO_EMUL
        n = 4;
#else
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
#endif
    }
    return n;
}
int warp_passes() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_XW"); v = e ? std::atoi(e) : 3; }
    return v;
}
bool size_supported(int n) {
    LG_SIZE_PAIRS(LG_SUP)
    return false;
}
}  // namespace lg
