// CPU unit test of the Stockham building blocks against a naive O(N^2) DFT.
#include <complex>
#include <cstdio>
#include <vector>

#include "../../lesgo_b200/csrc/fft_core.h"

using namespace lg;
typedef std::complex<double> cd;

template <int R> static double test_dft() {
    cplx v[R]; cd x[R];
    for (int i = 0; i < R; ++i) { x[i] = cd(std::sin(1.0 + i * 0.7), std::cos(0.3 * i * i)); v[i] = make_double2(x[i].real(), x[i].imag()); }
    Dft<R>::run(v);
    double err = 0;
    for (int k = 0; k < R; ++k) {
        cd s = 0;
        for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * ((k * n) % R) / R);
        err = std::max(err, std::abs(s - cd(v[k].x, v[k].y)));
    }
    return err;
}

template <int N, bool INV> __global__ void k_fft(const cplx* in, cplx* out, const cplx* W) {
    LG_DYN_SMEM(cplx, sm);
    constexpr int NF = 3;
    cplx* A = sm; cplx* B = sm + NF * SmemLen<N>::value;
    fft_tile<N, INV, NF, false, 64, false, false>(A, W,
        [](int f, int i) { return f * SmemLen<N>::value + spad(i); },
        [&](int f, int i) { return in[f * N + i]; },
        [&](int f, int i, cplx v) { out[f * N + i] = v; });
    // and the transform-fastest variant into the second half of out
    fft_tile<N, INV, NF, true, 64, false, false>(A, W,
        [](int f, int i) { return spad(i) * NF + f; },
        [&](int f, int i) { return in[f * N + i]; },
        [&](int f, int i, cplx v) { out[(NF + f) * N + i] = v; });
}

template <int N> static double test_fft() {
    constexpr int NF = 3;
    std::vector<cplx> in(NF * N), out(2 * NF * N), W(N);
    for (int m = 0; m < N; ++m) W[m] = make_double2(std::cos(2 * M_PI * m / N), -std::sin(2 * M_PI * m / N));
    for (int i = 0; i < NF * N; ++i) in[i] = make_double2(std::sin(0.37 * i + 1), std::cos(0.11 * i * 1.3));
    double err = 0;
    for (int inv = 0; inv < 2; ++inv) {
        const cplx* pi = in.data(); cplx* po = out.data(); const cplx* pw = W.data();
        size_t smem = 2 * NF * SmemLen<N>::value * sizeof(cplx);
        if (inv) { LG_LAUNCH((k_fft<N, true>), dim3(1), dim3(64), smem, 0, pi, po, pw); }
        else     { LG_LAUNCH((k_fft<N, false>), dim3(1), dim3(64), smem, 0, pi, po, pw); }
        for (int f = 0; f < NF; ++f)
            for (int k = 0; k < N; ++k) {
                cd s = 0;
                for (int n = 0; n < N; ++n)
                    s += cd(in[f * N + n].x, in[f * N + n].y) * std::polar(1.0, (inv ? 2 : -2) * M_PI * double((long)k * n % N) / N);
                err = std::max(err, std::abs(s - cd(out[f * N + k].x, out[f * N + k].y)) / N);
                err = std::max(err, std::abs(s - cd(out[(NF + f) * N + k].x, out[(NF + f) * N + k].y)) / N);
            }
    }
    return err;
}

int main() {
    int bad = 0;
#define TD(R) { double e = test_dft<R>(); printf("dft%-3d err %.2e\n", R, e); if (!(e < 1e-14)) ++bad; }
    TD(2) TD(3) TD(4) TD(5) TD(6) TD(8) TD(10) TD(12) TD(16)
#define TF(N) { double e = test_fft<N>(); printf("fft%-5d err %.2e\n", N, e); if (!(e < 1e-15)) ++bad; }
    TF(8) TF(12) TF(16) TF(24) TF(32) TF(48) TF(64) TF(80) TF(96) TF(128) TF(160) TF(192) TF(256) TF(320)
    TF(384) TF(512) TF(640) TF(768) TF(1024) TF(1536)
    printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
    return bad;
}
