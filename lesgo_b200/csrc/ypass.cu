#include "launch.h"
#include "sizes.h"
namespace lg {
template <int NIN, int NOUT>
static int launch_y_n(const YArgs& a, int nfields, int nplanes, const cplx* Win, const cplx* Wout,
                      cudaStream_t s) {
    typedef YCfg<NIN, NOUT> C;
    static bool attr = false;
    if (!attr) { set_smem(k_ypass<NIN, NOUT>, C::smem); attr = true; }
    if (nplanes <= 0 || nfields <= 0) return 0;
    dim3 grid((a.ncols + C::TC - 1) / C::TC, nplanes, nfields);
    LG_LAUNCH((k_ypass<NIN, NOUT>), grid, dim3(kBlock), C::smem, s, a, Win, Wout);
    return 0;
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
#define LG_SUP(S, B) if (n == S) return true;
bool size_supported(int n) {
    LG_SIZE_PAIRS(LG_SUP)
    return false;
}
}  // namespace lg
