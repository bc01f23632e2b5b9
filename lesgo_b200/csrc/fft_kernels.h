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
#include <type_traits>
#include "fft_core.h"

namespace lg {

constexpr int kMaxFields = 9;     // filt_da of u, v, w in one go: nine x-inverse outputs

// ---------------------------------------------------------------------------------
// x forward:  real rows -> half spectrum rows
// ---------------------------------------------------------------------------------
struct XfOut {
    double* dst[kMaxFields];   // per field: interleaved complex rows
    long plane;                // doubles between planes
    int row;                   // doubles between rows
    int ncol;                  // complex columns to write (<= nx/2); column ncol gets
    int write_nyq;             //   0: nothing, 1: zero, 2: the true Nyquist value
    int ring;                  // > 0: dst holds only `ring` planes, plane k lives in slot k % ring
};

// plane offset of a (possibly ring-buffered) intermediate array
LG_HD long poff(int k, long plane, int ring) { return long(ring > 0 ? k % ring : k) * plane; }

// spectral intermediates are read once and may have been written by another SM moments ago
// (plane pipeline, pipe_kernels.h): read them through L2 only
#ifdef LESGO_EMUL
LG_HD cplx ld_cg(const double* p) { return *reinterpret_cast<const cplx*>(p); }
#else
LG_D cplx ld_cg(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
#endif

template <int NX> struct XCfg {
    static constexpr int M = NX / 2;
    typedef TileGeom<M> G;
    // rows per tile: as many as a <= 256-thread block can take (see TileGeom)
    static constexpr int NF0 = 256 / G::threads(1);
    static constexpr int NF = NF0 < 1 ? 1 : (NF0 > 64 ? 64 : NF0);
    static constexpr int NTHR = G::round32(G::threads(NF));
    static constexpr int MINB = G::min_blocks(NTHR);
    static constexpr int SL = SmemLen<M>::value;
    static constexpr int NWH = M / 2 + 1;
    // work buffer + stage twiddles W_M + untangling twiddles W_NX[0..M/2]
    static constexpr int TWL = PlanInfo<M>::twlen;
    static constexpr size_t smem = size_t(NF * SL + TWL + NWH) * sizeof(cplx);
};

// A prologue may split load() into row() (per-row case analysis) and load_row() (straight-line element reads)
struct NoRow {};
template <class Pro, class = void> struct ProRow { typedef NoRow type; static constexpr bool value = false; };
template <class Pro> struct ProRow<Pro, std::void_t<typename Pro::Row>> { typedef typename Pro::Row type; static constexpr bool value = true; };

// one work item (tile of NF rows of one field) of the x-forward pass; every thread of the
// NTHR-thread block calls it; ends with a barrier
template <int NX, class Pro>
LG_D void xfwd_work(cplx* buf, const cplx* W, const cplx* Wh, int* s_k, int* s_y, double** s_p, const Pro& pro,
                    const XfOut& out, int nfields, int ny, int k0, int nplanes, unsigned work) {
    typedef XCfg<NX> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL, NTHR = C::NTHR;
    const unsigned nrows = unsigned(ny) * unsigned(nplanes);
    {
        // 32-bit index arithmetic throughout: a 64-bit division costs ~100 instructions per thread
        const int fld = int(work % unsigned(nfields));
        const unsigned row0 = (work / unsigned(nfields)) * NF;
        __shared__ typename ProRow<Pro>::type s_row[NF];
        for (int f = threadIdx.x; f < NF; f += NTHR) {
            const unsigned r = row0 + f;
            const int k = r < nrows ? k0 + int(r / unsigned(ny)) : -1, y = int(r % unsigned(ny));
            s_k[f] = k;
            s_y[f] = y;
            s_p[f] = out.dst[fld] + poff(k < 0 ? 0 : k, out.plane, out.ring) + long(y) * out.row;   // output row
            if constexpr (ProRow<Pro>::value) s_row[f] = pro.row(fld, k, y);
        }
        __syncthreads();
        if constexpr (ProRow<Pro>::value) {
            fft_tile<M, false, NF, false, NTHR, false, true, 1>(buf, W,
                [](int f) { return f * SL; },
                [&](int f, int i) { return pro.load_row(s_row[f], i); },
                [&](int f, int i, cplx v) { buf[f * SL + spad(i)] = v; });
        } else {
            fft_tile<M, false, NF, false, NTHR, false, true, 1>(buf, W,
                [](int f) { return f * SL; },
                [&](int f, int i) {
                    const int k = s_k[f];
                    if (k < 0) return make_double2(0.0, 0.0);
                    return pro.load(fld, k, s_y[f], i);
                },
                [&](int f, int i, cplx v) { buf[f * SL + spad(i)] = v; });
        }

        // untangle: X_k = E_k + W_N^k O_k,  X_{M-k} = conj(E_k - W_N^k O_k), pairs (m, M-m)
        constexpr int NP = M / 2 - 1;                       // pairs with 0 < m < M/2
        constexpr int ITER = (NF * NP + NTHR - 1) / NTHR;
#pragma unroll 2
        for (int q = 0; q < ITER; ++q) {
            const int it = threadIdx.x + q * NTHR;
            if (it >= NF * NP) break;
            const int m = 1 + it % NP, f = it / NP;
            const int k = s_k[f];
            if (k < 0) continue;
            double* drow = s_p[f];
            cplx a = buf[f * SL + spad(m)];
            cplx bz = buf[f * SL + spad(M - m)];
            cplx b = make_double2(bz.x, -bz.y);
            cplx e = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
            cplx d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
            cplx o = make_double2(d.y, -d.x);               // d / i
            cplx t = cmul(o, Wh[m]);
            if (m < out.ncol) *reinterpret_cast<cplx*>(drow + 2 * m) = cadd(e, t);
            if (M - m < out.ncol)
                *reinterpret_cast<cplx*>(drow + 2 * (M - m)) = make_double2(e.x - t.x, -(e.y - t.y));
        }
        // m = 0 (DC + Nyquist) and m = M/2 (self-paired): X_{M/2} = conj(Z_{M/2})
        for (int f = threadIdx.x; f < 2 * NF; f += NTHR) {
            const int ff = f >> 1, k = s_k[ff];
            if (k < 0) continue;
            double* drow = s_p[ff];
            if (f & 1) {
                cplx a = buf[ff * SL + spad(M / 2)];
                if (M / 2 < out.ncol) *reinterpret_cast<cplx*>(drow + M) = make_double2(a.x, -a.y);
            } else {
                cplx a = buf[ff * SL];
                if (out.ncol > 0) *reinterpret_cast<cplx*>(drow) = make_double2(a.x + a.y, 0.0);
                if (out.write_nyq && out.ncol >= M)
                    *reinterpret_cast<cplx*>(drow + 2 * M) = make_double2(out.write_nyq == 2 ? a.x - a.y : 0.0, 0.0);
                else if (out.write_nyq && out.ncol < M)
                    *reinterpret_cast<cplx*>(drow + 2 * out.ncol) = make_double2(0.0, 0.0);
            }
        }
        __syncthreads();
    }
}

// Pro: struct with  LG_D double2 load(int fld, int k, int y, int j) const
//      returning (x[2j], x[2j+1]) of row y of plane k of field fld (already scaled).
// Persistent blocks: each block loops over row tiles, twiddles live in shared memory.
// resident blocks asked of the compiler for prologues with a row plan (several operand loads per element): fewer
// blocks = more registers for loads in flight
#ifndef LG_XFV_MINB
#define LG_XFV_MINB 0
#endif
template <class Pro, class C> struct XFMinB {
    static constexpr int value = (ProRow<Pro>::value && LG_XFV_MINB > 0 && LG_XFV_MINB < C::MINB) ? LG_XFV_MINB : C::MINB;
};

template <int NX, class Pro>
__global__ void __launch_bounds__(XCfg<NX>::NTHR, XFMinB<Pro, XCfg<NX>>::value)
k_xfwd(const __grid_constant__ Pro pro, const __grid_constant__ XfOut out, int nfields, int zmajor, int ny, int k0, int nplanes,
       const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef XCfg<NX> C;
    constexpr int NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    cplx* buf = sm;
    cplx* W = sm + NF * SL;
    cplx* Wh = W + C::TWL;
    __shared__ int s_k[NF], s_y[NF];
    __shared__ double* s_p[NF];
    load_table(W, Wg, C::TWL);
    load_table(Wh, Whg, C::NWH);
    // Work items are dealt round-robin with the field index fastest: blocks resident together
    // work on neighbouring rows of all fields, so inputs the fields share (the u x omega
    // products) are read from DRAM once and concurrent blocks touch adjacent DRAM pages.
    // (Measured alternative, profiles/r1_work_order.md: a contiguous range per block, optionally
    // walking up z, was 25 % slower.)
    (void)zmajor;
    const long nrows = long(ny) * nplanes;
    const long nwork = ((nrows + NF - 1) / NF) * nfields;
    for (long work = blockIdx.x; work < nwork; work += gridDim.x)
        xfwd_work<NX, Pro>(buf, W, Wh, s_k, s_y, s_p, pro, out, nfields, ny, k0, nplanes, unsigned(work));
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
    int ring;      // > 0: src is a ring of `ring` planes
};

// one work item of the x-inverse pass (see xfwd_work)
template <int NX, class Epi>
LG_D void xinv_work(cplx* buf, const cplx* W, const cplx* Wh, int* s_k, int* s_y, double** s_p, const XiSrc& in,
                    const Epi& epi, int nfields, int ny, int k0, int nplanes, unsigned work) {
    typedef XCfg<NX> C;
    constexpr int M = C::M, NF = C::NF, SL = C::SL, NTHR = C::NTHR;
    const unsigned nrows = unsigned(ny) * unsigned(nplanes);
    {
        const int fld = int(work % unsigned(nfields));
        const unsigned row0 = (work / unsigned(nfields)) * NF;
        for (int f = threadIdx.x; f < NF; f += NTHR) {
            const unsigned r = row0 + f;
            const int k = r < nrows ? k0 + int(r / unsigned(ny)) : -1, y = int(r % unsigned(ny));
            s_k[f] = k;
            s_y[f] = y;
            s_p[f] = const_cast<double*>(in.src[fld]) + poff(k < 0 ? 0 : k, in.plane, in.ring) + long(y) * in.row;   // source row
        }
        __syncthreads();

        // tangle: Z'_k = E'_k + i O'_k,  E' = X_k + conj(X_{M-k}),  O' = (X_k - conj(X_{M-k})) conj(W_N^k)
        constexpr int NP = M / 2 - 1;
        constexpr int ITER = (NF * NP + NTHR - 1) / NTHR;
        constexpr int UN = 2;                                // global loads in flight per thread: 2*UN
#pragma unroll 1
        for (int q0 = 0; q0 < ITER; q0 += UN) {
            cplx va[UN], vb[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {                   // loads first
                const int it = threadIdx.x + (q0 + u) * NTHR;
                va[u] = make_double2(0.0, 0.0); vb[u] = va[u];
                if (it < NF * NP) {
                    const int m = 1 + it % NP, f = it / NP;
                    if (s_k[f] >= 0) {
                        const double* srow = s_p[f];
                        if (m < in.ncol) va[u] = ld_cg(srow + 2 * m);
                        if (M - m < in.ncol) vb[u] = ld_cg(srow + 2 * (M - m));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = threadIdx.x + (q0 + u) * NTHR;
                if (it < NF * NP) {
                    const int m = 1 + it % NP, f = it / NP;
                    cplx a = va[u], b = make_double2(vb[u].x, -vb[u].y);
                    cplx e = cadd(a, b);
                    cplx o = cmulc(csub(a, b), Wh[m]);
                    buf[f * SL + spad(m)] = make_double2(e.x - o.y, e.y + o.x);          // e + i o
                    buf[f * SL + spad(M - m)] = make_double2(e.x + o.y, -e.y + o.x);     // conj(e) + i conj(o)
                }
            }
        }
        for (int f = threadIdx.x; f < 2 * NF; f += NTHR) {
            const int ff = f >> 1;
            cplx z = make_double2(0.0, 0.0);
            if (s_k[ff] >= 0) {
                const double* srow = s_p[ff];
                if (f & 1) {                                  // m = M/2: Z' = 2 conj(X_{M/2})
                    if (M / 2 < in.ncol) { cplx a = ld_cg(srow + M); z = make_double2(2.0 * a.x, -2.0 * a.y); }
                } else {                                      // m = 0: real parts of X_0 and X_M only
                    double x0 = in.ncol > 0 ? ld_cg(srow).x : 0.0;
                    double xm = in.ncol > M ? ld_cg(srow + 2 * M).x : 0.0;
                    z = make_double2(x0 + xm, x0 - xm);
                }
            }
            buf[ff * SL + ((f & 1) ? spad(M / 2) : 0)] = z;
        }
        __syncthreads();

        fft_tile<M, true, NF, false, NTHR, true, false, 1>(buf, W,
            [](int f) { return f * SL; },
            [&](int f, int i) { return buf[f * SL + spad(i)]; },
            [&](int f, int i, cplx v) {
                if (s_k[f] < 0) return;
                epi.store(fld, s_k[f], s_y[f], i, v);
            });
        for (int f = threadIdx.x; f < NF; f += NTHR) {
            if (s_k[f] < 0) continue;
            epi.finish_row(fld, s_k[f], s_y[f]);
        }
        __syncthreads();
    }
}

// Epi: struct with  LG_D void store(int fld, int k, int y, int j, double2 v) const
//      receiving (x[2j], x[2j+1]);   LG_D void finish_row(int fld, int k, int y) const
template <int NX, class Epi>
__global__ void __launch_bounds__(XCfg<NX>::NTHR, XCfg<NX>::MINB)
k_xinv(const __grid_constant__ XiSrc in, const __grid_constant__ Epi epi, int nfields, int zmajor, int ny, int k0, int nplanes,
       const cplx* __restrict__ Wg, const cplx* __restrict__ Whg) {
    typedef XCfg<NX> C;
    constexpr int NF = C::NF, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    cplx* buf = sm;
    cplx* W = sm + NF * SL;
    cplx* Wh = W + C::TWL;
    __shared__ int s_k[NF], s_y[NF];
    __shared__ double* s_p[NF];
    load_table(W, Wg, C::TWL);
    load_table(Wh, Whg, C::NWH);
    (void)zmajor;   // round-robin work order, field index fastest: see k_xfwd
    const long nrows = long(ny) * nplanes;
    const long nwork = ((nrows + NF - 1) / NF) * nfields;
    for (long work = blockIdx.x; work < nwork; work += gridDim.x)
        xinv_work<NX, Epi>(buf, W, Wh, s_k, s_y, s_p, in, epi, nfields, ny, k0, nplanes, unsigned(work));
}

// ---------------------------------------------------------------------------------
// y pass
// ---------------------------------------------------------------------------------
enum YMode { Y_COPY = 0, Y_IKX = 1, Y_IKY = 2, Y_TABLE = 3, Y_TABLE2 = 4 };

struct YOutSpec {
    double* dst;
    int mode;
};
struct YField {
    const double* src;
    YOutSpec out[3];
    double* out2;            // NOUT2 kernels: also the 3/2-rule padded inverse transform of the spectrum (or null)
    // optional linear combination formed while loading (spectral reuse in lesgo_gpu_step, where the
    // vorticity's x spectra are combinations of x spectra filt_da already produced):
    //     in(k) = c0 * src(k) + c1 * (src2(k) - src2(k-1)) + c2 * src3(k)
    int combo;               // 0: in(k) = src(k)
    const double* src2;
    const double* src3;
    double c0, c1, c2;
};
struct YArgs {
    YField fld[kMaxFields];
    int nout;            // outputs per field (1..3)
    long src_plane, dst_plane;
    int src_row, dst_row;   // doubles
    int ncols;           // kx columns to transform
    int ncols2;          // > 0: Y_TABLE2 outputs only for column tiles that start below it (table2 is zero beyond)
    int k0, nplanes, nfields;
    double kxs, kys;     // 2 pi / L_x, 2 pi / L_y
    const double* table; // Y_TABLE multiplier, [ns][table_row] reals
    const double* table2; // Y_TABLE2 multiplier (second test filter of the same spectrum), same layout
    int table_row;
    int zero_col;        // >= 0: also write zeros into this complex column of every output row
    int keep_nyq_row;    // 1: raw transform (do not zero ky = ns/2)
    int src_ring, dst_ring;   // > 0: src / dst are rings of that many planes
    long dst2_plane;          // layout of the out2 arrays
    int dst2_row;
};

// MULTI = several outputs per field: the spectrum is kept in its own buffer S while the
// inverse transforms run in the work buffer; otherwise one buffer serves both.
// NOUT2 > 0 (with NIN == NOUT, MULTI): one more inverse transform of the SAME spectrum, padded to length
// NOUT2 -- lesgo_gpu_step's filt_da hands convec the 3/2-grid y transform of u, v, w directly, which
// saves convec's own y pass over those three fields (forward transform + read of the x spectra).
template <int NIN, int NOUT, bool MULTI, int NOUT2 = 0> struct YCfg {
    static_assert(NOUT2 == 0 || (MULTI && NIN == NOUT && NOUT2 > NOUT), "NOUT2 rides on the same-size multi-output pass");
    static constexpr int NMAX0 = NIN > NOUT ? NIN : NOUT;
    static constexpr int NMAX = NMAX0 > NOUT2 ? NMAX0 : NOUT2;
    static constexpr int NS = (NIN == 0) ? NOUT : ((NOUT == 0) ? NIN : (NIN < NOUT ? NIN : NOUT));  // spectral (small) length
    static constexpr int max2(int a, int b) { return a > b ? a : b; }
    // Two-stage column plans (fft_core.h Plan2) per kernel, by measurement on B200 at 512 x 512 x 256
    // (profiles/r4_experiments.md): they win for the forward-only pass and for filt_da's pass with the padded second
    // output, with 8 columns per tile; the pad / truncate passes keep the three-stage plans (the radix-32
    // butterflies cost 255 registers, which leaves one or two resident blocks).  LG_Y2_POLICY bits: 1 forward only,
    // 2 same-size + padded output, 4 inverse only with several outputs, 8 everything else.
#ifndef LG_Y2_POLICY
#define LG_Y2_POLICY 3
#endif
    static constexpr bool ALL2 = (NIN <= 0 || Plan2<(NIN > 0 ? NIN : 8)>::on) && (NOUT <= 0 || Plan2<(NOUT > 0 ? NOUT : 8)>::on) &&
                                 (NOUT2 <= 0 || Plan2<(NOUT2 > 0 ? NOUT2 : 8)>::on);
    static constexpr int KIND = (NOUT == 0) ? 1 : (NOUT2 > 0 ? 2 : ((NIN == 0 && MULTI) ? 4 : 8));
    static constexpr bool USE2 = ALL2 && (LG_Y2_POLICY & KIND) != 0;
    template <int N> static constexpr int thr(int tc) {
        return N <= 0 ? 0 : (USE2 ? tc * Plan2Info<(N > 0 ? N : 8)>::tpf : TileGeom<(N > 0 ? N : 8)>::threads(tc));
    }
    template <int N> static constexpr int twl() {
        return N <= 0 ? 0 : (USE2 ? Plan2Info<(N > 0 ? N : 8)>::twlen : PlanInfo<(N > 0 ? N : 8)>::twlen);
    }
    // offset of the table a kernel reads inside the per-length device table (Stockham stage tables first)
    static constexpr int TWOFF_IN = (NIN > 0 && USE2) ? PlanInfo<(NIN > 0 ? NIN : 8)>::twlen : 0;
    static constexpr int TWOFF_OUT = (NOUT > 0 && USE2) ? PlanInfo<(NOUT > 0 ? NOUT : 8)>::twlen : 0;
    static constexpr int TWOFF_OUT2 = (NOUT2 > 0 && USE2) ? PlanInfo<(NOUT2 > 0 ? NOUT2 : 8)>::twlen : 0;
    // 4 columns = 64 contiguous bytes per row; 8 columns (128 B) measured no faster (30.8 vs 30.0 ms/step)
#ifndef LG_Y12_TC2
#define LG_Y12_TC2 0
#endif
#ifndef LG_Y12_REGS
#define LG_Y12_REGS 112
#endif
    static constexpr int regs0 = max2(TileGeom<(NIN > 0 ? NIN : 8)>::regs * (NIN > 0), TileGeom<(NOUT > 0 ? NOUT : 8)>::regs * (NOUT > 0));
    // radix-12 plans (the 3/2-grid lengths) spill 300-700 bytes per thread under the 80-register cap;
    // -DLG_Y12_TC2=1: two columns per tile and LG_Y12_REGS registers.  Measured with 112 (spills remain)
    // and 160 registers (no spills, 12 warps/SM): pad 3.99 / 4.07 ms, trunc 1.73 / 1.80 ms against
    // 3.99 / 1.81 ms -- no difference, off by default
    static constexpr bool R12 = LG_Y12_TC2 && regs0 == 80;
#ifndef LG_Y2_TC
#define LG_Y2_TC 8
#endif
#ifndef LG_Y2_REGS
#define LG_Y2_REGS 200
#endif
    static constexpr int TC = USE2 ? LG_Y2_TC : ((max2(thr<NIN>(4), thr<NOUT>(4)) <= 512 && !R12) ? 4 : 2);   // (2 columns for the 768 passes: 30.8 ms)
    static constexpr int NTHR = ((max2(max2(thr<NIN>(TC), thr<NOUT>(TC)), USE2 ? thr<NOUT2>(TC) : 0) + 31) / 32) * 32;
    static constexpr int regs = USE2 ? LG_Y2_REGS : (R12 ? LG_Y12_REGS : regs0);
    static constexpr int SL = SmemLen<NMAX>::value;          // work buffer row count (padded)
    static constexpr int SLS = SmemLen<NS>::value;           // spectrum buffer (MULTI)
    static constexpr int NBUF = MULTI ? 2 : 1;
    static constexpr int BUFS = TC * SL + (MULTI ? TC * (NOUT2 > 0 ? SLS : SL) : 0);
    static constexpr int TWI = twl<NIN>();
    // with NOUT2 the forward and inverse tables of the same length are stored once (shared memory is tight)
    static constexpr int TWO = (NOUT2 > 0) ? 0 : twl<NOUT>();
    static constexpr int TWO2 = twl<NOUT2>();
    // PREF (-DLG_Y_PREF=1): the next tile is prefetched (cp.async) into a staging buffer while the current
    // one is being transformed, for the 3/2-rule pad passes (where the staging buffer does not cost a
    // resident block).  Measured: no gain (4.06 against 3.99 ms) -- the y passes are bound by the
    // wavefronts they push through the L1/shared-memory data pipe, not by load latency
    // (profiles/r2_experiments.md).  Off by default.
#ifndef LG_Y_PREF
#define LG_Y_PREF 0
#endif
    static constexpr bool PREF = LG_Y_PREF && NIN > 0 && NOUT > NIN && !MULTI && !USE2;
    static constexpr int STG = PREF ? NIN * TC : 0;
    static constexpr size_t smem = size_t(BUFS + TWI + TWO + TWO2 + STG) * sizeof(cplx);
    // resident blocks: the register budget is only capped as far as shared memory lets blocks fit
    static constexpr int by_regs = TileGeom<8>::blocks_for(NTHR, regs);
    static constexpr int by_smem = int((227 * 1024) / (smem + 1024)) < 1 ? 1 : int((227 * 1024) / (smem + 1024));
    static constexpr int MINB = by_regs < by_smem ? by_regs : by_smem;
};

// NIN  > 0: forward transform of length NIN first (input is the x-pass intermediate)
// NOUT > 0: inverse transform(s) of length NOUT last (output feeds the x inverse pass)
// NIN == NOUT: derivative / filter operators;  NIN < NOUT: padd;  NIN > NOUT: unpadd.
// Persistent blocks over (column tile, plane) pairs; blockIdx.y = field.
// element (row i, column f of the tile) of the y pass input, with the optional combination
LG_D cplx y_input(const YField& F, const YArgs& a, const double* src, int k, int c0col, int i, int f) {
    const long off = long(i) * a.src_row + 2 * f;
    cplx v = ld_cg(src + off);
    if (F.combo) {
        v = make_double2(F.c0 * v.x, F.c0 * v.y);
        if (F.src2) {
            const double* s2 = F.src2 + poff(k, a.src_plane, a.src_ring) + 2 * c0col + off;
            const double* s2m = F.src2 + poff(k - 1, a.src_plane, a.src_ring) + 2 * c0col + off;
            const cplx b = ld_cg(s2), bm = ld_cg(s2m);
            v.x += F.c1 * (b.x - bm.x);
            v.y += F.c1 * (b.y - bm.y);
        }
        if (F.src3) {
            const cplx d = ld_cg(F.src3 + poff(k, a.src_plane, a.src_ring) + 2 * c0col + off);
            v.x += F.c2 * d.x;
            v.y += F.c2 * d.y;
        }
    }
    return v;
}

// one work item (column tile of one plane of one field) of the y pass; shared memory is free
// again when it returns (its last transform ends with a barrier)
// start the asynchronous copy of work item `work`'s input tile (NIN rows x TC columns) into stg
template <int NIN, int NOUT, bool MULTI>
LG_D void ypass_prefetch(cplx* stg, const YArgs& a, unsigned work, unsigned nwork) {
    typedef YCfg<NIN, NOUT, MULTI> C;
    constexpr int TC = C::TC, NTHR = C::NTHR;
    if (work < nwork) {
        const int ntc = (a.ncols + TC - 1) / TC;
        const YField& F = a.fld[work % unsigned(a.nfields)];
        const unsigned tile = work / unsigned(a.nfields);
        const int c0 = int(tile % unsigned(ntc)) * TC;
        const int k = a.k0 + int(tile / unsigned(ntc));
        const int f = int(threadIdx.x) % TC;
        if (c0 + f < a.ncols) {
            const double* p = F.src + poff(k, a.src_plane, a.src_ring) + 2 * (c0 + f) + long(int(threadIdx.x) / TC) * a.src_row;
            const long step = long(NTHR / TC) * a.src_row;
#pragma unroll 4
            for (int it = threadIdx.x; it < NIN * TC; it += NTHR, p += step) cp_async16(stg + it, p);
        }
    }
    cp_async_commit();
}

// stg != nullptr: the input tile is already in the staging buffer (ypass_prefetch), and the tile of
// work item `next` is prefetched into it as soon as the first stage has read it
template <int NIN, int NOUT, bool MULTI, int NOUT2 = 0>
LG_D void ypass_work(cplx* buf, cplx* S, const cplx* Win, const cplx* Wout, const YArgs& a, unsigned work,
                     cplx* stg = nullptr, unsigned next = 0, unsigned nwork = 0, const cplx* Wout2 = nullptr) {
    typedef YCfg<NIN, NOUT, MULTI, NOUT2> C;
    constexpr int TC = C::TC, NS = C::NS, NTHR = C::NTHR;
    auto sidx = [](int f, int i) { return spad(i) * TC + f; };
    auto foff = [](int f) { return f; };
    const int ntc = (a.ncols + TC - 1) / TC;
    {
        const YField& F = a.fld[work % unsigned(a.nfields)];
        const unsigned tile = work / unsigned(a.nfields);
        const int c0 = int(tile % unsigned(ntc)) * TC;
        const int k = a.k0 + int(tile / unsigned(ntc));
        const double* src = F.src + poff(k, a.src_plane, a.src_ring) + 2 * c0;
        // every item of a thread lies in the same tile column (NTHR is a multiple of TC): its column
        // checks and i*kx factor are per-tile constants
        const int f_t = int(threadIdx.x) % TC;
        const bool colok = c0 + f_t < a.ncols;
        const double kx = a.kxs * double(c0 + f_t);

        if constexpr (NIN > 0) {
            if constexpr (NOUT == 0) {
                // forward only: straight to global with Nyquist-row zeroing
                double* dst = F.out[0].dst + poff(k, a.dst_plane, a.dst_ring) + 2 * c0;
                fft_tile_cols<NIN, false, TC, NTHR, false, false, TC, C::USE2>(buf, Win, foff,
                    [&](int f, int i) {
                        if (!colok) return make_double2(0.0, 0.0);
                        return y_input(F, a, src, k, c0, i, f);
                    },
                    [&](int f, int i, cplx v) {
                        if (!colok) return;
                        if (i == NIN / 2 && !a.keep_nyq_row) v = make_double2(0.0, 0.0);
                        *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * f) = v;
                    });
            } else {
                auto keep = [&](int f, int i, cplx v) {
                    // keep only the NS rows of the small spectrum (unpadd, fft.f90:86-97)
                    int is = i;
                    if (NIN > NS) {
                        if (i < NS / 2) is = i;
                        else if (i > NIN - NS / 2) is = i - (NIN - NS);
                        else return;
                    }
                    S[sidx(f, is)] = v;
                };
                if (C::PREF && stg) {
                    const double sc = F.combo ? F.c0 : 1.0;
                    fft_tile_cols<NIN, false, TC, NTHR, false, !MULTI, TC, C::USE2>(buf, Win, foff,
                        [&](int f, int i) {
                            if (!colok) return make_double2(0.0, 0.0);
                            const cplx v = stg[i * TC + f];
                            return make_double2(sc * v.x, sc * v.y);
                        },
                        keep, [&]() { ypass_prefetch<NIN, NOUT, MULTI>(stg, a, next, nwork); });
                } else {
                    fft_tile_cols<NIN, false, TC, NTHR, false, !MULTI, TC, C::USE2>(buf, Win, foff,
                        [&](int f, int i) {
                            if (!colok) return make_double2(0.0, 0.0);
                            return y_input(F, a, src, k, c0, i, f);
                        },
                        keep);
                }
            }
        } else {
#pragma unroll 8
            for (int it = threadIdx.x; it < TC * NS; it += NTHR) {
                int f = it % TC, i = it / TC;
                cplx v = make_double2(0.0, 0.0);
                if (c0 + f < a.ncols) v = ld_cg(src + long(i) * a.src_row + 2 * f);
                S[sidx(f, i)] = v;
            }
            __syncthreads();
        }

        if constexpr (NOUT > 0) {
            // i*kx does not depend on y, so the y-inverse of (i kx S) is i kx times the y-inverse
            // of S: an IKX output is derived from the COPY transform instead of costing its own.
            int o_copy = -1, o_ikx = -1;
            for (int o = 0; o < a.nout; ++o) {
                if (F.out[o].mode == Y_COPY) o_copy = o;
                if (F.out[o].mode == Y_IKX) o_ikx = o;
            }
            for (int o = 0; o < a.nout; ++o) {
                const int mode = F.out[o].mode;
                if (mode == Y_IKX && o_copy >= 0) continue;          // written by the COPY transform
                if (mode == Y_TABLE2 && a.ncols2 > 0 && c0 >= a.ncols2) continue;   // nobody reads those columns
                double* dst = F.out[o].dst + poff(k, a.dst_plane, a.dst_ring) + 2 * c0;
                double* dst_x = (mode == Y_COPY && o_ikx >= 0) ? F.out[o_ikx].dst + poff(k, a.dst_plane, a.dst_ring) + 2 * c0 : nullptr;
                fft_tile_cols<NOUT, true, TC, NTHR, !MULTI, false, TC, C::USE2>(buf, Wout, foff,
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
                        if (mode == Y_COPY || mode == Y_IKX) return v;
                        if (mode == Y_IKY) {
                            double ky = a.kys * double(is < NS / 2 ? is : is - NS);
                            return make_double2(-v.y * ky, v.x * ky);
                        }
                        const double* tb = mode == Y_TABLE2 ? a.table2 : a.table;
                        double g = colok ? tb[long(is) * a.table_row + c0 + f] : 0.0;
                        return make_double2(v.x * g, v.y * g);
                    },
                    [&](int f, int i, cplx v) {
                        if (!colok) return;
                        if (mode == Y_IKX) v = make_double2(-v.y * kx, v.x * kx);
                        *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * f) = v;
                        if (dst_x) *reinterpret_cast<cplx*>(dst_x + long(i) * a.dst_row + 2 * f) = make_double2(-v.y * kx, v.x * kx);
                    });
            }
        }
        if constexpr (NOUT2 > 0) {
            if (F.out2) {
                // padd (fft.f90:60-69) + inverse transform of length NOUT2 of the same spectrum
                double* dst2 = F.out2 + long(k) * a.dst2_plane + 2 * c0;
                fft_tile_cols<NOUT2, true, TC, NTHR, false, false, TC, C::USE2>(buf, Wout2, foff,
                    [&](int f, int i) {
                        int is;
                        if (i < NS / 2) is = i;
                        else if (i > NOUT2 - NS / 2) is = i - (NOUT2 - NS);
                        else return make_double2(0.0, 0.0);
                        return S[sidx(f, is)];
                    },
                    [&](int f, int i, cplx v) {
                        if (!colok) return;
                        *reinterpret_cast<cplx*>(dst2 + long(i) * a.dst2_row + 2 * f) = v;
                    });
            }
        }
        if (a.zero_col >= 0 && c0 == 0) {
            constexpr int NR = NOUT > 0 ? NOUT : NIN;
            for (int o = 0; o < a.nout; ++o) {
                double* dst = F.out[o].dst + poff(k, a.dst_plane, a.dst_ring);
                for (int i = threadIdx.x; i < NR; i += NTHR)
                    *reinterpret_cast<cplx*>(dst + long(i) * a.dst_row + 2 * a.zero_col) = make_double2(0.0, 0.0);
            }
        }
    }
}

template <int NIN, int NOUT, bool MULTI, int NOUT2 = 0>
__global__ void __launch_bounds__(YCfg<NIN, NOUT, MULTI, NOUT2>::NTHR, YCfg<NIN, NOUT, MULTI, NOUT2>::MINB)
k_ypass(const __grid_constant__ YArgs a, const cplx* __restrict__ Wing, const cplx* __restrict__ Woutg,
        const cplx* __restrict__ Wout2g = nullptr) {
    typedef YCfg<NIN, NOUT, MULTI, NOUT2> C;
    constexpr int TC = C::TC, SL = C::SL;
    LG_DYN_SMEM(cplx, sm);
    cplx* buf = sm;                                     // work buffer
    cplx* S = MULTI ? sm + TC * SL : sm;                // spectrum of the tile (NS rows used)
    cplx* Win = sm + C::BUFS;
    cplx* Wout = NOUT2 > 0 ? Win : Win + C::TWI;        // NOUT2: same length, same table
    cplx* Wout2 = Win + C::TWI + C::TWO;
    if (NIN > 0) load_table(Win, Wing + C::TWOFF_IN, C::TWI);
    if (NOUT > 0 && NOUT2 == 0) load_table(Wout, Woutg + C::TWOFF_OUT, C::TWO);
    if (NOUT2 > 0) load_table(Wout2, Wout2g + C::TWOFF_OUT2, C::TWO2);
    __syncthreads();
    const int ntc = (a.ncols + TC - 1) / TC;
    const long nwork = long(ntc) * a.nplanes * a.nfields;
    if constexpr (C::PREF) {
        // fields that need more than a scaled copy on load (y_input) take the direct path
        bool plain = true;
        for (int i = 0; i < a.nfields; ++i) plain = plain && !a.fld[i].src2 && !a.fld[i].src3;
        if (plain) {
            cplx* stg = Wout2 + C::TWO2;
            ypass_prefetch<NIN, NOUT, MULTI>(stg, a, blockIdx.x, unsigned(nwork));
            for (long work = blockIdx.x; work < nwork; work += gridDim.x) {
                cp_async_wait_all();
                __syncthreads();
                ypass_work<NIN, NOUT, MULTI>(buf, S, Win, Wout, a, unsigned(work), stg, unsigned(work + gridDim.x), unsigned(nwork));
            }
            return;
        }
    }
    for (long work = blockIdx.x; work < nwork; work += gridDim.x)    // round-robin: see k_xfwd
        ypass_work<NIN, NOUT, MULTI, NOUT2>(buf, S, Win, Wout, a, unsigned(work), nullptr, 0, 0, Wout2);
}

}  // namespace lg
