// fft_core.h -- FP64 radix-2/3/4/5/6/8/10/12/16 Stockham building blocks.
//
// Replaces the FFTW plans of the reference (fft.f90:114-121).  Everything here is
// host/device so the same code runs on sm_100a and in the CPU kernel-logic emulator.
//
// A length-N complex FFT is a sequence of up to four Stockham autosort stages with
// radices R1*R2*R3*R4 = N.  Stage s (radix R, Ns = product of earlier radices) has
// N/R work items j:   v[r] = in[j + r*N/R] * W_N^{(j % Ns) * r * N/(Ns*R)}
//                      v    = DFT_R(v)
//                      out[(j/Ns)*Ns*R + (j%Ns) + r*Ns] = v[r]
// so reads are always unit-stride in j (coalesced from global memory in the first
// stage, conflict-free from shared memory later), the output of the last stage is in
// natural order, and its writes are again unit-stride in j.  The inverse transform
// uses the same code on (im, re)-swapped data, i.e. conjugated twiddles.
#pragma once
#include "portable.h"

#ifndef LG_TW_PRODUCT
#define LG_TW_PRODUCT 1
#endif

namespace lg {

typedef double2 cplx;

LG_HD cplx cmul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
LG_HD cplx cmulc(cplx a, cplx b) {   // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
LG_HD cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
LG_HD cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
LG_HD cplx cswap(cplx a) { return make_double2(a.y, a.x); }
LG_HD cplx cmul_mi(cplx a) { return make_double2(a.y, -a.x); }   // a * (-i)
LG_HD cplx cmul_pi(cplx a) { return make_double2(-a.y, a.x); }   // a * (+i)
LG_HD cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }

// ---------------------------------------------------------------------------------
// Forward (sign -1) small DFTs on registers.
// ---------------------------------------------------------------------------------
template <int R> struct Dft;

template <> struct Dft<1> { static LG_HD void run(cplx*) {} };

template <> struct Dft<2> {
    static LG_HD void run(cplx* v) {
        cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <> struct Dft<3> {
    static LG_HD void run(cplx* v) {
        const double s = 0.86602540378443864676;   // sqrt(3)/2
        cplx t1 = cadd(v[1], v[2]);
        cplx t2 = make_double2(fma(-0.5, t1.x, v[0].x), fma(-0.5, t1.y, v[0].y));
        cplx d = csub(v[1], v[2]);
        cplx t3 = make_double2(s * d.y, -s * d.x);   // -i * s * d
        v[0] = cadd(v[0], t1);
        v[1] = cadd(t2, t3);
        v[2] = csub(t2, t3);
    }
};

template <> struct Dft<4> {
    static LG_HD void run(cplx* v) {
        cplx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        cplx t2 = cadd(v[1], v[3]), t3 = cmul_mi(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <> struct Dft<5> {
    static LG_HD void run(cplx* v) {
        const double c1 = 0.30901699437494742410;    // cos(2pi/5)
        const double c2 = -0.80901699437494742410;   // cos(4pi/5)
        const double s1 = 0.95105651629515357212;    // sin(2pi/5)
        const double s2 = 0.58778525229247312917;    // sin(4pi/5)
        cplx a14 = cadd(v[1], v[4]), d14 = csub(v[1], v[4]);
        cplx a23 = cadd(v[2], v[3]), d23 = csub(v[2], v[3]);
        cplx x0 = v[0];
        cplx p1 = make_double2(fma(c1, a14.x, fma(c2, a23.x, x0.x)), fma(c1, a14.y, fma(c2, a23.y, x0.y)));
        cplx p2 = make_double2(fma(c2, a14.x, fma(c1, a23.x, x0.x)), fma(c2, a14.y, fma(c1, a23.y, x0.y)));
        // q = -i * (s1*d14 + s2*d23), r = -i * (s2*d14 - s1*d23)
        cplx u1 = make_double2(fma(s1, d14.x, s2 * d23.x), fma(s1, d14.y, s2 * d23.y));
        cplx u2 = make_double2(fma(s2, d14.x, -s1 * d23.x), fma(s2, d14.y, -s1 * d23.y));
        cplx q1 = cmul_mi(u1), q2 = cmul_mi(u2);
        v[0] = cadd(x0, cadd(a14, a23));
        v[1] = cadd(p1, q1);
        v[4] = csub(p1, q1);
        v[2] = cadd(p2, q2);
        v[3] = csub(p2, q2);
    }
};

template <> struct Dft<8> {
    static LG_HD void run(cplx* v) {
        const double h = 0.70710678118654752440;
        // even / odd 4-point DFTs (decimation in time)
        cplx e[4] = {v[0], v[2], v[4], v[6]};
        cplx o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        // twiddles W8^k, k=0..3: 1, (1-i)h, -i, (-1-i)h
        cplx o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        cplx o2 = cmul_mi(o[2]);
        cplx o3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
    }
};

template <> struct Dft<16> {
    static LG_HD void run(cplx* v) {
        // 4 x 4 Cooley-Tukey: n = 4*n1 + n2, k = k1 + 4*k2
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos,sin(pi/8)
        const double h = 0.70710678118654752440;
        cplx t[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            cplx u[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
            Dft<4>::run(u);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) t[n2][k1] = u[k1];
        }
        // twiddle W16^{n2*k1}
        t[1][1] = cmul(t[1][1], make_double2(c1, -s1));
        t[1][2] = cmul(t[1][2], make_double2(h, -h));
        t[1][3] = cmul(t[1][3], make_double2(s1, -c1));
        t[2][1] = cmul(t[2][1], make_double2(h, -h));
        t[2][2] = cmul_mi(t[2][2]);
        t[2][3] = cmul(t[2][3], make_double2(-h, -h));
        t[3][1] = cmul(t[3][1], make_double2(s1, -c1));
        t[3][2] = cmul(t[3][2], make_double2(-h, -h));
        t[3][3] = cmul(t[3][3], make_double2(-c1, s1));
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            cplx u[4] = {t[0][k1], t[1][k1], t[2][k1], t[3][k1]};
            Dft<4>::run(u);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = u[k2];
        }
    }
};

// 32 = 2 x 16, decimation in time: X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k]
template <> struct Dft<32> {
    static LG_HD void run(cplx* v) {
        cplx e[16], o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<16>::run(e);
        Dft<16>::run(o);
        const double h = 0.70710678118654752440;
        // cos, sin (2 pi k / 32), k = 1, 2, 3, 5, 6, 7 (k = 4 is h, h)
        const double c1 = 0.98078528040323044913, s1 = 0.19509032201612826785;
        const double c2 = 0.92387953251128675613, s2 = 0.38268343236508977173;
        const double c3 = 0.83146961230254523708, s3 = 0.55557023301960222474;
#define LG_BF32(k, t)  { const cplx t_ = (t); v[k] = cadd(e[k], t_); v[(k) + 16] = csub(e[k], t_); }
#define LG_TW32(a, c, s) make_double2(fma((a).x, (c), (a).y * (s)), fma((a).y, (c), -(a).x * (s)))   /* a * (c - i s) */
        LG_BF32(0, o[0])
        LG_BF32(1, LG_TW32(o[1], c1, s1))
        LG_BF32(2, LG_TW32(o[2], c2, s2))
        LG_BF32(3, LG_TW32(o[3], c3, s3))
        LG_BF32(4, make_double2(h * (o[4].x + o[4].y), h * (o[4].y - o[4].x)))
        LG_BF32(5, LG_TW32(o[5], s3, c3))
        LG_BF32(6, LG_TW32(o[6], s2, c2))
        LG_BF32(7, LG_TW32(o[7], s1, c1))
        LG_BF32(8, cmul_mi(o[8]))
        LG_BF32(9, LG_TW32(o[9], -s1, c1))
        LG_BF32(10, LG_TW32(o[10], -s2, c2))
        LG_BF32(11, LG_TW32(o[11], -s3, c3))
        LG_BF32(12, make_double2(h * (o[12].y - o[12].x), -h * (o[12].x + o[12].y)))
        LG_BF32(13, LG_TW32(o[13], -c3, s3))
        LG_BF32(14, LG_TW32(o[14], -c2, s2))
        LG_BF32(15, LG_TW32(o[15], -c1, s1))
#undef LG_BF32
#undef LG_TW32
    }
};

// Good-Thomas prime-factor composition for coprime N1, N2 (no twiddles):
// input n = (N2*n1 + N1*n2) mod N, output k with k = k1 (mod N1), k = k2 (mod N2).
template <int N1, int N2> struct DftPfa {
    static constexpr int N = N1 * N2;
    static LG_HD void run(cplx* v) {
        cplx t[N2][N1];
#pragma unroll
        for (int n2 = 0; n2 < N2; ++n2) {
            cplx u[N1];
#pragma unroll
            for (int n1 = 0; n1 < N1; ++n1) u[n1] = v[(N2 * n1 + N1 * n2) % N];
            Dft<N1>::run(u);
#pragma unroll
            for (int k1 = 0; k1 < N1; ++k1) t[n2][k1] = u[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) {
            cplx u[N2];
#pragma unroll
            for (int n2 = 0; n2 < N2; ++n2) u[n2] = t[n2][k1];
            Dft<N2>::run(u);
#pragma unroll
            for (int k2 = 0; k2 < N2; ++k2) v[crt_table(k1, k2)] = u[k2];
        }
    }
    static LG_HD int crt_table(int k1, int k2) {
        // k = k1*a + k2*b mod N with a = N2*(N2^-1 mod N1), b = N1*(N1^-1 mod N2)
        return (k1 * A + k2 * B) % N;
    }
    static constexpr int inv_mod(int a, int m) {
        for (int x = 1; x < m; ++x)
            if ((a * x) % m == 1) return x;
        return 1;
    }
    static constexpr int A = N2 * inv_mod(N2 % N1, N1);
    static constexpr int B = N1 * inv_mod(N1 % N2, N2);
};
template <> struct Dft<6> { static LG_HD void run(cplx* v) { DftPfa<2, 3>::run(v); } };
template <> struct Dft<10> { static LG_HD void run(cplx* v) { DftPfa<2, 5>::run(v); } };
template <> struct Dft<12> { static LG_HD void run(cplx* v) { DftPfa<4, 3>::run(v); } };
template <> struct Dft<24> { static LG_HD void run(cplx* v) { DftPfa<8, 3>::run(v); } };

// ---------------------------------------------------------------------------------
// Plans.  LG_PLAN(N, R1, R2, R3, R4): radices multiply to N; unused stages are 1.
// Radix order and the 64-register cap were chosen by measurement (profiles/r1_plan_variants.md).
// ---------------------------------------------------------------------------------
template <int N> struct Plan;
#define LG_PLAN(N_, A_, B_, C_, D_)                                                  \
    template <> struct Plan<N_> {                                                    \
        static constexpr int N = N_, R1 = A_, R2 = B_, R3 = C_, R4 = D_;             \
        static_assert(A_ * B_ * C_ * D_ == N_, "radices must multiply to N");        \
    };
LG_PLAN(8, 8, 1, 1, 1)
LG_PLAN(12, 12, 1, 1, 1)
LG_PLAN(16, 4, 4, 1, 1)
LG_PLAN(24, 6, 4, 1, 1)
LG_PLAN(32, 8, 4, 1, 1)
LG_PLAN(36, 6, 6, 1, 1)
LG_PLAN(40, 10, 4, 1, 1)
LG_PLAN(48, 6, 8, 1, 1)
LG_PLAN(60, 10, 6, 1, 1)
LG_PLAN(64, 8, 8, 1, 1)
LG_PLAN(72, 12, 6, 1, 1)
LG_PLAN(80, 10, 8, 1, 1)
LG_PLAN(96, 12, 8, 1, 1)
LG_PLAN(120, 6, 5, 4, 1)
LG_PLAN(128, 8, 4, 4, 1)
LG_PLAN(144, 6, 6, 4, 1)
LG_PLAN(160, 10, 4, 4, 1)
LG_PLAN(192, 6, 8, 4, 1)
LG_PLAN(240, 10, 6, 4, 1)
LG_PLAN(256, 8, 8, 4, 1)
LG_PLAN(288, 6, 6, 8, 1)
LG_PLAN(320, 10, 8, 4, 1)
LG_PLAN(384, 8, 8, 6, 1)
LG_PLAN(480, 10, 8, 6, 1)
LG_PLAN(512, 8, 8, 8, 1)
LG_PLAN(576, 12, 6, 8, 1)
LG_PLAN(640, 10, 8, 8, 1)
LG_PLAN(768, 8, 8, 12, 1)
LG_PLAN(1024, 8, 4, 4, 8)
LG_PLAN(1536, 6, 8, 8, 4)
#undef LG_PLAN

// max work items of any stage / min: threads per FFT are sized to the widest radix
template <int N> struct PlanInfo {
    typedef Plan<N> P;
    static constexpr int rmax = (P::R1 > P::R2 ? P::R1 : P::R2) > (P::R3 > P::R4 ? P::R3 : P::R4)
                                    ? (P::R1 > P::R2 ? P::R1 : P::R2)
                                    : (P::R3 > P::R4 ? P::R3 : P::R4);
    static constexpr int rmin_nz(int a, int b) { return b == 1 ? a : (a < b ? a : b); }
    static constexpr int rmin = rmin_nz(rmin_nz(rmin_nz(P::R1, P::R2), P::R3), P::R4);
    static constexpr int threads = N / rmax;       // threads cooperating on one FFT
    static constexpr int nstages = 1 + (P::R2 > 1) + (P::R3 > 1) + (P::R4 > 1);
    // per-stage twiddle tables, stage s >= 2: (R_s - 1) * Ns_s entries, concatenated
    static constexpr int off2 = 0;
    static constexpr int off3 = off2 + (P::R2 > 1 ? (P::R2 - 1) * P::R1 : 0);
    static constexpr int off4 = off3 + (P::R3 > 1 ? (P::R3 - 1) * P::R1 * P::R2 : 0);
    static constexpr int twlen = off4 + (P::R4 > 1 ? (P::R4 - 1) * P::R1 * P::R2 * P::R3 : 0);
};

// ---------------------------------------------------------------------------------
// One Stockham stage for work item j (0 <= j < N/R), split into its load+butterfly half
// and its store half so a block can run the stage IN PLACE on one shared buffer (all
// loads, barrier, all stores).
//   INV = false: forward (sign -1); true: inverse (sign +1, unnormalised).
//   ld(idx) -> cplx   : fetch logical element idx of the stage input
//   st(idx, cplx)     : deliver logical element idx of the stage output
//   Wst               : this stage's twiddles, Wst[(r-1)*Ns + k] (shared memory)
// ---------------------------------------------------------------------------------
template <int N, int R, int Ns, bool INV, class Ld>
LG_HD void stage_load(int j, const cplx* __restrict__ Wst, Ld ld, cplx* v) {
    constexpr int T = N / R;
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ld(j + r * T);
    if (Ns > 1) {
        // Wst[(r-1)*Ns + k] = W_N^{k r N/(Ns R)}: lanes with consecutive k read consecutive
        // entries (conflict-free), lanes with equal k broadcast
        const int k = j % Ns;
#if LG_TW_PRODUCT
        // The block-cooperative passes are bound by shared-memory bandwidth, not by the FP64 pipe
        // (profiles/r2_experiments.md): fetch only w^1, w^2, w^4, w^8 and form the other powers as
        // products of those (at most two multiplications deep, so no error growth along a chain).
        // only p1, p2, p4, p8 and w^3 stay live (20 registers), each power is applied as soon as formed
        static_assert(R <= 12, "twiddle products are written for radices up to 12");
        const cplx p1 = Wst[k];
        const cplx p2 = R > 2 ? Wst[Ns + k] : p1;
        const cplx p4 = R > 4 ? Wst[3 * Ns + k] : p1;
        const cplx p8 = R > 8 ? Wst[7 * Ns + k] : p1;
        const cplx w3 = cmul(p2, p1);
#pragma unroll
        for (int r = 1; r < R; ++r) {
            const cplx hi = r >= 8 ? p8 : (r >= 4 ? p4 : (r >= 2 ? p2 : p1));
            const int lo = r - (r >= 8 ? 8 : (r >= 4 ? 4 : (r >= 2 ? 2 : 1)));
            const cplx w = lo == 0 ? hi : cmul(hi, lo == 1 ? p1 : (lo == 2 ? p2 : w3));
            v[r] = INV ? cmulc(v[r], w) : cmul(v[r], w);
        }
#else
#pragma unroll
        for (int r = 1; r < R; ++r) {
            cplx w = Wst[(r - 1) * Ns + k];
            v[r] = INV ? cmulc(v[r], w) : cmul(v[r], w);
        }
#endif
    }
    if (INV) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
    }
    Dft<R>::run(v);
    if (INV) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
    }
}

template <int N, int R, int Ns, class St>
LG_HD void stage_store(int j, St st, const cplx* v) {
    const int j0 = (Ns == 1) ? j * R : ((j / Ns) * Ns * R + (j % Ns));
#pragma unroll
    for (int r = 0; r < R; ++r) st(j0 + r * Ns, v[r]);
}

// Shared-memory index padding: one extra cplx every 8 (keeps the radix-strided
// writes of the first stage conflict-free for 16-byte accesses).
LG_HD int spad(int i) { return i + (i >> 3); }
template <int N> struct SmemLen { static constexpr int value = N + (N >> 3) + 1; };

constexpr int kBlock = 256;     // threads per block of the pointwise kernels

// ---- block geometry of a tile of NF length-N transforms -----------------------------------
// A thread never holds more than RCAP = (largest radix of the plan) complex values: in a
// stage of radix R it may take floor(RCAP/R) butterflies.  tile_threads<N>(NF) is the block
// size that lets every stage finish in one such pass.
template <int N> struct TileGeom {
    typedef Plan<N> P;
#ifndef LG_RCAP_MUL
#define LG_RCAP_MUL 1
#endif
    static constexpr int RCAP = PlanInfo<N>::rmax * LG_RCAP_MUL;
    static constexpr int need(int nf, int r) { return r == 1 ? 0 : (nf * (N / r) + (RCAP / r) - 1) / (RCAP / r); }
    static constexpr int max2(int a, int b) { return a > b ? a : b; }
    static constexpr int threads(int nf) {
        return max2(max2(need(nf, P::R1), need(nf, P::R2)), max2(need(nf, P::R3), need(nf, P::R4)));
    }
    static constexpr int round32(int t) { return ((t + 31) / 32) * 32; }
    // registers per thread the butterflies want, and the resident blocks that allows
    static constexpr bool pow2(int r) { return r == 1 || r == 2 || r == 4 || r == 8 || r == 16; }
    static constexpr bool pure2 = pow2(P::R1) && pow2(P::R2) && pow2(P::R3) && pow2(P::R4);
    static constexpr int regs = RCAP <= 8 ? 64 : (RCAP <= 12 ? 80 : 128);   // 51 (5 blocks/SM) measured slower
    // (LG_RCAP_MUL = 2, two butterflies per thread and half the threads: measured in profiles/r2_experiments.md)
    static constexpr int min_blocks(int nthr) { return blocks_for(nthr, regs); }
    // at least two resident blocks for blocks of <= 512 threads, at most 16
    static constexpr int blocks_for(int nthr, int r) {
        return (65536 / (nthr * r)) < (nthr <= 512 ? 2 : 1) ? (nthr <= 512 ? 2 : 1)
                                                           : ((65536 / (nthr * r)) > 16 ? 16 : (65536 / (nthr * r)));
    }
};

// ---------------------------------------------------------------------------------
// Block-cooperative FFT over a tile of NF independent length-N transforms, in place on
// ONE shared buffer.
//   FFT_FASTEST = true : consecutive threads take consecutive transforms (column tiles
//                        of the y pass: the transform index is the contiguous global
//                        direction);  false: consecutive threads take consecutive
//                        butterflies of one transform (row tiles of the x pass).
//   ld(f, i)      : element i of transform f for the first stage
//   st(f, i, v)   : element i of the result (natural order) from the last stage
//   LD_BUF/ST_BUF : ld reads / st writes the work buffer itself (adds the barriers an
//                   in-place update then needs)
//   sidx(f, i)    : shared-memory slot of element i of transform f
// Every thread of the NTHR-thread block must call this; it ends with a barrier.
// ---------------------------------------------------------------------------------
// Shared-memory slot of element i of transform f: (spad(i) * ES + foff(f)), with ES = 1,
// foff = f*SL for row tiles and ES = TC, foff = f for column tiles.  When the stride between
// the R operands of a butterfly is a multiple of 8 the padded slots are an arithmetic
// progression, so the stage computes one base address and uses constant offsets.
template <int N, int R, int Ns, bool INV, int NF, bool FFT_FASTEST, int NTHR, bool SRC_SMEM, bool DST_SMEM,
          bool SYNC_BETWEEN, int ES, class FOff, class Ld, class St>
LG_D void tile_stage(cplx* buf, const cplx* __restrict__ W, FOff foff, Ld ld, St st) {
    constexpr int T = N / R, ITEMS = NF * T, IPT = (ITEMS + NTHR - 1) / NTHR;
    cplx v[IPT][R];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int it = threadIdx.x + q * NTHR;
        if (ITEMS % NTHR == 0 || it < ITEMS) {
            int f, j;
            if (FFT_FASTEST) { f = it % NF; j = it / NF; }
            else             { j = it % T;  f = it / T; }
            if constexpr (SRC_SMEM) {
                if constexpr (T % 8 == 0) {
                    const cplx* p = buf + spad(j) * ES + foff(f);
                    stage_load<N, R, Ns, INV>(j, W, [&](int i) { return p[((i - j) + ((i - j) >> 3)) * ES]; }, v[q]);
                } else {
                    const int fo = foff(f);
                    stage_load<N, R, Ns, INV>(j, W, [&](int i) { return buf[spad(i) * ES + fo]; }, v[q]);
                }
            } else {
                stage_load<N, R, Ns, INV>(j, W, [&](int i) { return ld(f, i); }, v[q]);
            }
        }
    }
    if (SYNC_BETWEEN) __syncthreads();
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int it = threadIdx.x + q * NTHR;
        if (ITEMS % NTHR == 0 || it < ITEMS) {
            int f, j;
            if (FFT_FASTEST) { f = it % NF; j = it / NF; }
            else             { j = it % T;  f = it / T; }
            if constexpr (DST_SMEM) {
                const int j0 = (Ns == 1) ? j * R : ((j / Ns) * Ns * R + (j % Ns));
                if constexpr (Ns % 8 == 0 || (Ns == 1 && (R == 8 || R == 4 || R == 2))) {
                    // j0 + r*Ns: (j0 + r*Ns) >> 3 == (j0 >> 3) + r*Ns/8 in both cases
                    cplx* p = buf + spad(j0) * ES + foff(f);
                    constexpr int step = (Ns % 8 == 0) ? (Ns + Ns / 8) : 1;
#pragma unroll
                    for (int r = 0; r < R; ++r) p[r * step * ES] = v[q][r];
                } else {
                    const int fo = foff(f);
#pragma unroll
                    for (int r = 0; r < R; ++r) buf[spad(j0 + r * Ns) * ES + fo] = v[q][r];
                }
            } else {
                stage_store<N, R, Ns>(j, [&](int i, cplx x) { st(f, i, x); }, v[q]);
            }
        }
    }
}

struct NoHook { LG_HD void operator()() const {} };

// hook(): called by every thread once the first stage has consumed its input (after the barrier
// that follows it), e.g. to start the asynchronous prefetch of the next tile into the staging
// buffer the first stage just read
template <int N, bool INV, int NF, bool FFT_FASTEST, int NTHR, bool LD_BUF, bool ST_BUF, int ES, class FOff, class Ld, class St,
          class Hook = NoHook>
LG_D void fft_tile(cplx* buf, const cplx* __restrict__ W, FOff foff, Ld ld, St st, Hook hook = Hook()) {
    typedef Plan<N> P;
    typedef PlanInfo<N> PI;
    constexpr int NST = PI::nstages;
    if constexpr (NST == 1) {
        tile_stage<N, P::R1, 1, INV, NF, FFT_FASTEST, NTHR, false, false, LD_BUF && ST_BUF, ES>(buf, W, foff, ld, st);
    } else if constexpr (NST == 2) {
        tile_stage<N, P::R1, 1, INV, NF, FFT_FASTEST, NTHR, false, true, LD_BUF, ES>(buf, W, foff, ld, st);
        __syncthreads();
        hook();
        tile_stage<N, P::R2, P::R1, INV, NF, FFT_FASTEST, NTHR, true, false, ST_BUF, ES>(buf, W + PI::off2, foff, ld, st);
    } else if constexpr (NST == 3) {
        tile_stage<N, P::R1, 1, INV, NF, FFT_FASTEST, NTHR, false, true, LD_BUF, ES>(buf, W, foff, ld, st);
        __syncthreads();
        hook();
        tile_stage<N, P::R2, P::R1, INV, NF, FFT_FASTEST, NTHR, true, true, true, ES>(buf, W + PI::off2, foff, ld, st);
        __syncthreads();
        tile_stage<N, P::R3, P::R1 * P::R2, INV, NF, FFT_FASTEST, NTHR, true, false, ST_BUF, ES>(buf, W + PI::off3, foff, ld, st);
    } else {
        tile_stage<N, P::R1, 1, INV, NF, FFT_FASTEST, NTHR, false, true, LD_BUF, ES>(buf, W, foff, ld, st);
        __syncthreads();
        hook();
        tile_stage<N, P::R2, P::R1, INV, NF, FFT_FASTEST, NTHR, true, true, true, ES>(buf, W + PI::off2, foff, ld, st);
        __syncthreads();
        tile_stage<N, P::R3, P::R1 * P::R2, INV, NF, FFT_FASTEST, NTHR, true, true, true, ES>(buf, W + PI::off3, foff, ld, st);
        __syncthreads();
        tile_stage<N, P::R4, P::R1 * P::R2 * P::R3, INV, NF, FFT_FASTEST, NTHR, true, false, ST_BUF, ES>(buf, W + PI::off4, foff, ld, st);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// Two-stage plans N = R1 * R2 with register butterflies of radix 16 ... 32 (column transforms of the y pass).
// A column makes ONE round trip through shared memory instead of the two of a three-stage radix-8 plan, with
// one block barrier inside the transform instead of two, and about half the address arithmetic per point: the
// y passes are bound by the instructions and shared-memory wavefronts they issue per point, not by arithmetic
// (profiles/r4_experiments.md).
//   stage 1 (radix R1, no twiddles):  work item j < N/R1 = R2 reads elements j + r*R2, writes slot R1*j + r
//   stage 2 (radix R2, Ns = R1):      work item j < N/R2 = R1 reads slots j + r*R1 ( * W_N^{j r}), writes element j + r*R1
// Shared-memory slot s of a column lives at (s + s/R1) * ES + f: the stage-1 stores of neighbouring work items are
// (R1+1) elements apart (odd -> the two halves of a 128-byte wavefront), the stage-2 loads are contiguous in j.
// Twiddles: W_N^{j r} = T_hi[r & ~7][j] * T_lo[r & 7][j] from a table of 7 + (R2-1)/8 rows of R1 entries (one
// multiplication deep).
// ---------------------------------------------------------------------------------
#ifndef LG_Y2STAGE
#define LG_Y2STAGE 1
#endif
template <int N> struct Plan2 { static constexpr bool on = false; static constexpr int R1 = 1, R2 = 1; };
#if LG_Y2STAGE
template <> struct Plan2<256> { static constexpr bool on = true; static constexpr int R1 = 16, R2 = 16; };
template <> struct Plan2<384> { static constexpr bool on = true; static constexpr int R1 = 16, R2 = 24; };
template <> struct Plan2<512> { static constexpr bool on = true; static constexpr int R1 = 16, R2 = 32; };
template <> struct Plan2<768> { static constexpr bool on = true; static constexpr int R1 = 32, R2 = 24; };
#endif
template <int N> struct Plan2Info {
    typedef Plan2<N> P;
    static constexpr int RMAX = P::R1 > P::R2 ? P::R1 : P::R2;
    static constexpr int T1 = N / P::R1, T2 = N / P::R2;
    static constexpr int per1 = RMAX / P::R1;                            // stage-1 butterflies a thread may hold
    static constexpr int tpf1 = (T1 + per1 - 1) / per1;
    static constexpr int tpf = tpf1 > T2 ? tpf1 : T2;                    // threads per transform
    static constexpr int nrows = 7 + (P::R2 - 1) / 8;
    static constexpr int twlen = nrows * P::R1;
    static constexpr int slots = N + N / P::R1 + 1;                      // padded column length
};

// F_FAST: consecutive threads take consecutive transforms (column tiles); else consecutive butterflies of one
// transform (row tiles, ES = 1: keeps the stage-1 stores of a warp (R1+1) slots apart = conflict-free)
template <int N, bool INV, int NF, int NTHR, bool LD_BUF, bool ST_BUF, int ES, bool F_FAST = true, class FOff, class Ld, class St,
          class Hook = NoHook>
LG_D void fft_tile2(cplx* buf, const cplx* __restrict__ W, FOff foff, Ld ld, St st, Hook hook = Hook()) {
    typedef Plan2<N> P;
    constexpr int R1 = P::R1, R2 = P::R2, T1 = N / R1, T2 = N / R2;
    static_assert(T1 == R2 && T2 == R1, "two-stage plan");
    {
        constexpr int ITEMS = NF * T1, IPT = (ITEMS + NTHR - 1) / NTHR;
        cplx v[IPT][R1];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int it = threadIdx.x + q * NTHR;
            if (ITEMS % NTHR == 0 || it < ITEMS) {
                const int f = F_FAST ? it % NF : it / (ITEMS / NF), j = F_FAST ? it / NF : it % (ITEMS / NF);
#pragma unroll
                for (int r = 0; r < R1; ++r) v[q][r] = ld(f, j + r * T1);
                if (INV) {
#pragma unroll
                    for (int r = 0; r < R1; ++r) v[q][r] = cswap(v[q][r]);
                }
                Dft<R1>::run(v[q]);
                if (INV) {
#pragma unroll
                    for (int r = 0; r < R1; ++r) v[q][r] = cswap(v[q][r]);
                }
            }
        }
        if (LD_BUF) __syncthreads();
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int it = threadIdx.x + q * NTHR;
            if (ITEMS % NTHR == 0 || it < ITEMS) {
                const int f = F_FAST ? it % NF : it / (ITEMS / NF), j = F_FAST ? it / NF : it % (ITEMS / NF);
                cplx* p = buf + (R1 + 1) * j * ES + foff(f);
#pragma unroll
                for (int r = 0; r < R1; ++r) p[r * ES] = v[q][r];
            }
        }
    }
    __syncthreads();
    hook();
    {
        constexpr int ITEMS = NF * T2, IPT = (ITEMS + NTHR - 1) / NTHR;
        cplx v[IPT][R2];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int it = threadIdx.x + q * NTHR;
            if (ITEMS % NTHR == 0 || it < ITEMS) {
                const int f = F_FAST ? it % NF : it / (ITEMS / NF), j = F_FAST ? it / NF : it % (ITEMS / NF);
                const cplx* p = buf + j * ES + foff(f);
#pragma unroll
                for (int r = 0; r < R2; ++r) v[q][r] = p[r * (R1 + 1) * ES];
                // twiddles W_N^{j r}: rows 0..6 hold r = 1..7, rows 7.. hold r = 8, 16, 24
                cplx lo[8];
#pragma unroll
                for (int m = 1; m < 8 && m < R2; ++m) lo[m] = W[(m - 1) * R1 + j];
#pragma unroll
                for (int hb = 0; hb < R2; hb += 8) {
                    cplx hi = make_double2(1.0, 0.0);
                    if (hb > 0) hi = W[(6 + hb / 8) * R1 + j];
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int r = hb + m;
                        if (r == 0 || r >= R2) continue;
                        const cplx w = hb == 0 ? lo[m] : (m == 0 ? hi : cmul(hi, lo[m]));
                        v[q][r] = INV ? cmulc(v[q][r], w) : cmul(v[q][r], w);
                    }
                }
                if (INV) {
#pragma unroll
                    for (int r = 0; r < R2; ++r) v[q][r] = cswap(v[q][r]);
                }
                Dft<R2>::run(v[q]);
                if (INV) {
#pragma unroll
                    for (int r = 0; r < R2; ++r) v[q][r] = cswap(v[q][r]);
                }
            }
        }
        if (ST_BUF) __syncthreads();
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int it = threadIdx.x + q * NTHR;
            if (ITEMS % NTHR == 0 || it < ITEMS) {
                const int f = F_FAST ? it % NF : it / (ITEMS / NF), j = F_FAST ? it / NF : it % (ITEMS / NF);
#pragma unroll
                for (int r = 0; r < R2; ++r) st(f, j + r * R1, v[q][r]);
            }
        }
    }
    __syncthreads();
}

// column tiles of the y pass: the two-stage plan (USE2, chosen per kernel by YCfg) or the Stockham stages above
template <int N, bool INV, int NF, int NTHR, bool LD_BUF, bool ST_BUF, int ES, bool USE2, class FOff, class Ld, class St,
          class Hook = NoHook>
LG_D void fft_tile_cols(cplx* buf, const cplx* __restrict__ W, FOff foff, Ld ld, St st, Hook hook = Hook()) {
    if constexpr (USE2) fft_tile2<N, INV, NF, NTHR, LD_BUF, ST_BUF, ES, true>(buf, W, foff, ld, st, hook);
    else fft_tile<N, INV, NF, true, NTHR, LD_BUF, ST_BUF, ES>(buf, W, foff, ld, st, hook);
}

// 16-byte asynchronous global -> shared copies (LDGSTS): prefetch of the next tile while the
// current one is being transformed
#ifdef LESGO_EMUL
LG_HD void cp_async16(cplx* dst, const double* src) { *dst = *reinterpret_cast<const cplx*>(src); }
LG_HD void cp_async_commit() {}
LG_HD void cp_async_wait_all() {}
#else
LG_D void cp_async16(cplx* dst, const double* src) {
    const unsigned d = unsigned(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
LG_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
LG_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// Bulk asynchronous copies (the TMA engine's 1-D form, cp.async.bulk -> UBLKCP) of whole contiguous rows into
// shared memory, completion counted in bytes on an mbarrier: one thread issues, everybody waits on the phase parity.
// src, dst and bytes must be multiples of 16.
#ifndef LESGO_EMUL
LG_D void mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = unsigned(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
LG_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = unsigned(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
LG_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    const unsigned d = unsigned(__cvta_generic_to_shared(dst)), a = unsigned(__cvta_generic_to_shared(bar));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(a) : "memory");
}
LG_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = unsigned(__cvta_generic_to_shared(bar));
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
#endif

// length of the concatenated stage-twiddle table of Plan<N> and its (R1..R4) for the host
struct PlanDesc { int n, r[4], twlen; };
template <int N> inline PlanDesc plan_desc() {
    PlanDesc d; d.n = N; d.r[0] = Plan<N>::R1; d.r[1] = Plan<N>::R2; d.r[2] = Plan<N>::R3; d.r[3] = Plan<N>::R4;
    d.twlen = PlanInfo<N>::twlen; return d;
}

// copy a twiddle table into shared memory (all threads; caller synchronises)
LG_D void load_table(cplx* dst, const cplx* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace lg
