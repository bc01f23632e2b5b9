// pipe.cu -- launchers of the plane-pipeline kernels (pipe_kernels.h), instantiated for the
// production grid lengths only; other lengths fall back to one launch per pass.
#include <cstdlib>
#include "launch.h"
#include "pipe_kernels.h"
namespace lg {

bool pipe_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_PIPE"); v = (e && e[0] == '1') ? 1 : 0; }   // opt-in: measured slower, profiles/r2_experiments.md
    return v != 0;
}
int pipe_ring() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_PIPE_RING"); v = e ? std::atoi(e) : 4; if (v < 1) v = 1; }
    return v;
}

template <int NXF, class Pro, int NIN, int NOUT, bool MULTI, int NXI, class Epi>
static int launch_pipe_t(const Pro& pro, const XfOut& xo, const YArgs& ya, const XiSrc& xi, const Epi& epi,
                         const PipeCtl& ctl, const cplx* Wx, const cplx* Whx, const cplx* Win, const cplx* Wout,
                         cudaStream_t s) {
    typedef PipeCfg<NXF, NIN, NOUT, MULTI, NXI> C;
    if constexpr (!C::ok) {
        return -1;
    } else {
        typedef typename C::CX CX;
        if ((C::HF && ctl.ny_f % CX::NF) || (C::HI && ctl.ny_i % CX::NF)) return -1;
        static bool attr = false;
        if (!attr) { set_smem(k_pipe<NXF, Pro, NIN, NOUT, MULTI, NXI, Epi>, C::smem); attr = true; }
        if (ctl.nplanes <= 0) return 0;
        const long tiles = long(ctl.nplanes) *
                           ((C::HF ? (ctl.ny_f / CX::NF) * ctl.nf_f : 0) + ((ya.ncols + C::CY::TC - 1) / C::CY::TC) * ya.nfields +
                            (C::HI ? (ctl.ny_i / CX::NF) * ctl.nf_i : 0));
        dim3 grid(persistent_blocks(C::smem, tiles, C::MINB));
        LG_LAUNCH((k_pipe<NXF, Pro, NIN, NOUT, MULTI, NXI, Epi>), grid, dim3(C::NTHR), C::smem, s, pro, xo, ya, xi, epi, ctl,
                  Wx, Whx, Win, Wout);
        return 0;
    }
}

// x-forward -> y (derivative operators) -> x-inverse: derivatives.f90 ddx / ddy / ddxy / filt_da
int launch_pipe_deriv(int nx, int ny, bool multi, const ProScale& pro, const XfOut& xo, const YArgs& ya,
                      const XiSrc& xi, const EpiStore& epi, const PipeCtl& ctl, const cplx* Wx, const cplx* Whx,
                      const cplx* Wy, cudaStream_t s) {
#define LG_PIPE_DERIV(NX, NY)                                                                              \
    if (nx == NX && ny == NY)                                                                              \
        return multi ? launch_pipe_t<NX, ProScale, NY, NY, true, NX, EpiStore>(pro, xo, ya, xi, epi, ctl, Wx, Whx, Wy, Wy, s) \
                     : launch_pipe_t<NX, ProScale, NY, NY, false, NX, EpiStore>(pro, xo, ya, xi, epi, ctl, Wx, Whx, Wy, Wy, s);
    LG_PIPE_SIZES(LG_PIPE_DERIV)
    return -1;
}

}  // namespace lg
