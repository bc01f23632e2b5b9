// fft_kernels.h -- batched 2-D real FFT passes with fused prologues / spectral
// operators / epilogues.  Together they replace every dfftw_execute_dft_r2c/c2r call
// site on the path and the padd/unpadd
// copies of fft.f90:43-99.
//
// A 2-D real transform of an (nx, ny) plane is two passes:
//   x pass (k_xfwd / k_xinv): one real row = one half-length complex FFT (nx/2 points,
//       packed z_j = x_2j + i x_2j+1) plus the real<->half-complex untangling step,
//       staged in shared memory; rows are contiguous so all global traffic is
//       unit-stride 16-byte accesses.
//   y pass (k_ypass): complex FFTs down columns of the half spectrum, TC adjacent
//       kx columns per block (16*TC contiguous bytes per row), forward transform,
//       spectral operator (i*kx, i*ky, filter table, 3/2-rule pad or truncate, Nyquist
//       zeroing) and up to three inverse transforms without leaving shared memory.
// Column kx = nx/2 (the "oddball", derivatives.f90:194) is never transformed: every
// consumer on this path zeroes it, so the y pass works on nx/2 columns and writes 0
// there.  Transforms are unnormalised like FFTW's.
#pragma once
#include "fft_core.h"

namespace lg {

constexpr int kMaxFields = 6;
constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------
// x forward:  real rows -> half spectrum rows
// ---------------------------------------------------------------------------------
struct XfOut {
    double* dst[kMaxFields];   // per field: interleaved complex rows
    long plane;                // doubles between planes
    int row;                   // doubles between rows
    int ncol;                  // complex columns to write (<= nx/2); column ncol gets
    int write_nyq;             //   0: nothing, 1: zero, 2: the true Nyquist value
};

template <int NX> struct XCfg {
    static constexpr int M = NX / 2;
    static constexpr int NF = (2048 / M) < 1 ? 1 : ((2048 / M) > 64 ? 64 : (2048 / M));
    static constexpr int SL = SmemLen<M>::value;
    static constexpr size_t smem = size_t(2) * NF * SL * sizeof(cplx);
};

// Pro: struct with   int nfields;   LG_D double2 load(int fld, int k, int y, int j) const
//      returning (x[2j], x[2j+1]) of row y of plane k of field fld (already scaled).
template <int NX, class Pro>
__global__ void __launch_bounds__(kBlock)
k_xfwd(const __grid_constant__ Pro pro, const __grid_constant__ XfOut out, int ny, int k0, int nplanes,
       const cplx* __restrict__ W, const cplx* __restrict__ Wh) {
    typedef XCfg<NX> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    cplx* A = sm;
    cplx* B = sm + NF * SL;
    const int fld = blockIdx.y;
    const long nrows = long(ny) * nplanes;
    const long row0 = long(blockIdx.x) * NF;
    constexpr int NST = PlanInfo<M>::nstages;
    cplx* Z = (NST & 1) ? A : B;   // buffer free to take the last stage's output

    fft_tile<M, false, NF, false>(A, B, W,
        [](int f, int i) { return f * SL + spad(i); },
        [&](int f, int i) {
            long r = row0 + f;
            if (r >= nrows) return make_double2(0.0, 0.0);
            int k = k0 + int(r / ny), y = int(r % ny);
            return pro.load(fld, k, y, i);
        },
        [&](int f, int i, cplx v) { Z[f * SL + spad(i)] = v; });

    // untangle: X_k = E_k + W_N^k O_k,  X_{M-k} = conj(E_k - W_N^k O_k)
    for (int it = threadIdx.x; it < NF * (M / 2 + 1); it += blockDim.x) {
        int m = it % (M / 2 + 1), f = it / (M / 2 + 1);
        long r = row0 + f;
        if (r >= nrows) continue;
        int k = k0 + int(r / ny), y = int(r % ny);
        double* drow = out.dst[fld] + long(k) * out.plane + long(y) * out.row;
        cplx a = Z[f * SL + spad(m)];
        if (m == 0) {
            if (out.ncol > 0) *reinterpret_cast<cplx*>(drow) = make_double2(a.x + a.y, 0.0);
            if (out.write_nyq && out.ncol >= M)
                *reinterpret_cast<cplx*>(drow + 2 * M) =
                    make_double2(out.write_nyq == 2 ? a.x - a.y : 0.0, 0.0);
            else if (out.write_nyq && out.ncol < M)
                *reinterpret_cast<cplx*>(drow + 2 * out.ncol) = make_double2(0.0, 0.0);
            continue;
        }
        cplx bz = Z[f * SL + spad(M - m)];
        cplx b = make_double2(bz.x, -bz.y);
        cplx e = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
        cplx d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
        cplx o = make_double2(d.y, -d.x);            // d / i
        cplx t = cmul(o, Wh[m]);
        if (m < out.ncol) *reinterpret_cast<cplx*>(drow + 2 * m) = cadd(e, t);
        if (m != M - m && M - m < out.ncol)
            *reinterpret_cast<cplx*>(drow + 2 * (M - m)) = make_double2(e.x - t.x, -(e.y - t.y));
    }
}

// ---------------------------------------------------------------------------------
// x inverse:  half spectrum rows -> real rows
// ---------------------------------------------------------------------------------
struct XiSrc {
    const double* src[kMaxFields];
    long plane;
    int row;
    int ncol;      // stored complex columns carrying data; columns >= ncol read as 0.
                   // ncol > nx/2 means the Nyquist column is present too.
};

// Epi: struct with  LG_D void store(int fld, int k, int y, int j, double2 v) const
//      receiving (x[2j], x[2j+1]);   LG_D void finish_row(int fld, int k, int y) const
template <int NX, class Epi>
__global__ void __launch_bounds__(kBlock)
k_xinv(const __grid_constant__ XiSrc in, const __grid_constant__ Epi epi, int ny, int k0, int nplanes,
       const cplx* __restrict__ W, const cplx* __restrict__ Wh) {
    typedef XCfg<NX> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    cplx* A = sm;
    cplx* B = sm + NF * SL;
    const int fld = blockIdx.y;
    const long nrows = long(ny) * nplanes;
    const long row0 = long(blockIdx.x) * NF;

    // tangle: Z'_k = E'_k + i O'_k,  E' = X_k + conj(X_{M-k}),  O' = (X_k - conj(X_{M-k})) conj(W_N^k)
    for (int it = threadIdx.x; it < NF * (M / 2 + 1); it += blockDim.x) {
        int m = it % (M / 2 + 1), f = it / (M / 2 + 1);
        long r = row0 + f;
        cplx za = make_double2(0.0, 0.0), zb = za;
        if (r < nrows) {
            int k = k0 + int(r / ny), y = int(r % ny);
            const double* srow = in.src[fld] + long(k) * in.plane + long(y) * in.row;
            if (m == 0) {
                double x0 = in.ncol > 0 ? srow[0] : 0.0;
                double xm = in.ncol > M ? srow[2 * M] : 0.0;
                za = make_double2(x0 + xm, x0 - xm);
            } else {
                cplx a = m < in.ncol ? *reinterpret_cast<const cplx*>(srow + 2 * m) : make_double2(0.0, 0.0);
                cplx bz = (M - m) < in.ncol ? *reinterpret_cast<const cplx*>(srow + 2 * (M - m)) : make_double2(0.0, 0.0);
                cplx b = make_double2(bz.x, -bz.y);
                cplx e = cadd(a, b);
                cplx o = cmulc(csub(a, b), Wh[m]);
                za = make_double2(e.x - o.y, e.y + o.x);         // e + i o
                zb = make_double2(e.x + o.y, -e.y + o.x);        // conj(e) + i conj(o)
            }
        }
        B[f * SL + spad(m)] = za;
        if (m != 0 && m != M - m) B[f * SL + spad(M - m)] = zb;
    }
    __syncthreads();

    fft_tile<M, true, NF, false>(A, B, W,
        [](int f, int i) { return f * SL + spad(i); },
        [&](int f, int i) { return B[f * SL + spad(i)]; },
        [&](int f, int i, cplx v) {
            long r = row0 + f;
            if (r >= nrows) return;
            int k = k0 + int(r / ny), y = int(r % ny);
            epi.store(fld, k, y, i, v);
        });
    for (int f = threadIdx.x; f < NF; f += blockDim.x) {
        long r = row0 + f;
        if (r >= nrows) continue;
        epi.finish_row(fld, k0 + int(r / ny), int(r % ny));
    }
}

// ---------------------------------------------------------------------------------
// y pass
// ---------------------------------------------------------------------------------
enum YMode { Y_COPY = 0, Y_IKX = 1, Y_IKY = 2, Y_TABLE = 3 };

struct YOutSpec {
    double* dst;
    int mode;
};
struct YField {
    const double* src;
    YOutSpec out[3];
};
struct YArgs {
    YField fld[kMaxFields];
    int nout;            // outputs per field (1..3)
    long src_plane, dst_plane;
    int src_row, dst_row;   // doubles
    int ncols;           // kx columns to transform
    int k0;
    double kxs, kys;     // 2 pi / L_x, 2 pi / L_y
    const double* table; // Y_TABLE multiplier, [ns][table_row] reals
    int table_row;
    int zero_col;        // >= 0: also write zeros into this complex column of every output row
    int keep_nyq_row;    // 1: raw transform (do not zero ky = ns/2)
};

template <int NIN, int NOUT> struct YCfg {
    static constexpr int NMAX = NIN > NOUT ? NIN : NOUT;
    static constexpr int NS = (NIN == 0) ? NOUT : ((NOUT == 0) ? NIN : (NIN < NOUT ? NIN : NOUT));  // spectral (small) length
    static constexpr int TC = NMAX <= 512 ? 4 : 2;
    static constexpr int SL = SmemLen<NMAX>::value;
    static constexpr int NBUF = (NIN > 0 && NOUT == 0) ? 2 : 3;
    static constexpr size_t smem = size_t(NBUF) * TC * SL * sizeof(cplx);
};

// NIN  > 0: forward transform of length NIN first (input is the x-pass intermediate)
// NOUT > 0: inverse transform(s) of length NOUT last (output feeds the x inverse pass)
// NIN == NOUT: derivative / filter operators;  NIN < NOUT: padd;  NIN > NOUT: unpadd.
template <int NIN, int NOUT>
__global__ void __launch_bounds__(kBlock)
k_ypass(const __grid_constant__ YArgs a, const cplx* __restrict__ Win, const cplx* __restrict__ Wout) {
    typedef YCfg<NIN, NOUT> C;
    constexpr int TC = C::TC, SL = C::SL, NS = C::NS;
    LG_DYN_SMEM(cplx, sm);
    cplx* S = sm;                                   // spectrum of the tile (NS rows used)
    cplx* A = sm + TC * SL;
    cplx* B = (C::NBUF == 3) ? sm + 2 * TC * SL : sm;
    const int c0 = blockIdx.x * TC;
    const int k = a.k0 + blockIdx.y;
    const YField& F = a.fld[blockIdx.z];
    auto sidx = [](int f, int i) { return spad(i) * TC + f; };
    const double* src = F.src + long(k) * a.src_plane + 2 * c0;

    if constexpr (NIN > 0) {
        if constexpr (NOUT == 0) {
            // forward only: straight to global with Nyquist-row zeroing
            double* dst = F.out[0].dst + long(k) * a.dst_plane + 2 * c0;
            fft_tile<NIN, false, TC, true>(S, A, Win, sidx,
                [&](int f, int i) {
                    if (c0 + f >= a.ncols) return make_double2(0.0, 0.0);
                    return *reinterpret_cast<const cplx*>(src + long(i) * a.src_row + 2 * f);
                },
                [&](int f, int i, cplx v) {
                    if (c0 + f >= a.ncols) return;
                    if (i == NIN / 2 && !a.keep_nyq_row) v = make_double2(0.0, 0.0);
                    *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * f) = v;
                });
        } else {
            fft_tile<NIN, false, TC, true>(A, B, Win, sidx,
                [&](int f, int i) {
                    if (c0 + f >= a.ncols) return make_double2(0.0, 0.0);
                    return *reinterpret_cast<const cplx*>(src + long(i) * a.src_row + 2 * f);
                },
                [&](int f, int i, cplx v) {
                    // keep only the NS rows of the small spectrum (unpadd, fft.f90:86-97)
                    int is = i;
                    if (NIN > NS) {
                        if (i < NS / 2) is = i;
                        else if (i > NIN - NS / 2) is = i - (NIN - NS);
                        else return;
                    }
                    S[sidx(f, is)] = v;
                });
        }
    } else {
        for (int it = threadIdx.x; it < TC * NS; it += blockDim.x) {
            int f = it % TC, i = it / TC;
            cplx v = make_double2(0.0, 0.0);
            if (c0 + f < a.ncols) v = *reinterpret_cast<const cplx*>(src + long(i) * a.src_row + 2 * f);
            S[sidx(f, i)] = v;
        }
        __syncthreads();
    }

    if constexpr (NOUT > 0) {
        for (int o = 0; o < a.nout; ++o) {
            const int mode = F.out[o].mode;
            double* dst = F.out[o].dst + long(k) * a.dst_plane + 2 * c0;
            fft_tile<NOUT, true, TC, true>(A, B, Wout, sidx,
                [&](int f, int i) {
                    // row i of the (possibly padded) output spectrum <- small row is
                    int is = i;
                    if (NOUT > NS) {                         // padd, fft.f90:60-69
                        if (i < NS / 2) is = i;
                        else if (i > NOUT - NS / 2) is = i - (NOUT - NS);
                        else return make_double2(0.0, 0.0);
                    }
                    if (is == NS / 2 && !a.keep_nyq_row) return make_double2(0.0, 0.0);
                    cplx v = S[sidx(f, is)];
                    if (mode == Y_COPY) return v;
                    if (mode == Y_IKX) {
                        double kx = a.kxs * double(c0 + f);
                        return make_double2(-v.y * kx, v.x * kx);
                    }
                    if (mode == Y_IKY) {
                        double ky = a.kys * double(is < NS / 2 ? is : is - NS);
                        return make_double2(-v.y * ky, v.x * ky);
                    }
                    double g = (c0 + f < a.ncols) ? a.table[long(is) * a.table_row + c0 + f] : 0.0;
                    return make_double2(v.x * g, v.y * g);
                },
                [&](int f, int i, cplx v) {
                    if (c0 + f >= a.ncols) return;
                    *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * f) = v;
                });
        }
    }
    if (a.zero_col >= 0 && blockIdx.x == 0) {
        constexpr int NR = NOUT > 0 ? NOUT : NIN;
        for (int o = 0; o < a.nout; ++o) {
            double* dst = F.out[o].dst + long(k) * a.dst_plane;
            for (int i = threadIdx.x; i < NR; i += blockDim.x)
                *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * a.zero_col) = make_double2(0.0, 0.0);
        }
    }
}

}  // namespace lg
