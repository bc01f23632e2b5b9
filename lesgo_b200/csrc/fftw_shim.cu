// fftw_shim.cu -- the FFTW3 legacy-Fortran entry points LESGO calls outside the hot path
// (SURVEY 8(b) "second, lower boundary"):
//     dfftw_plan_dft_r2c_2d_, dfftw_plan_dft_c2r_2d_, dfftw_execute_dft_r2c_, dfftw_execute_dft_c2r_,
//     dfftw_destroy_plan_
// as gfortran / ifort mangle them, all arguments by reference, the plan an integer*8, dimensions in
// Fortran order (n_fast, n_slow).  Callers in the reference: fft.f90:114-121 (the four plans of module
// fft), test_filtermodule.f90:138-167, scalars.f90:513-624, turbine_indicator.f90:130-151 (its own
// 2048 x 2048 out-of-place plans).  With these exported a LESGO build links without libfftw3.
//
// A plan is a slot in a small table.  Plans whose shape is the bound context's (nx, ny) or
// (3nx/2, 3ny/2) AND that were made in place run on the hot path's own kernels
// (lesgo_gpu_fft_r2c / _c2r); every other shape whose lengths factor into 2, 3, 5 runs on a generic
// global-memory Stockham transform below (start-up work such as the disk indicator convolution, not
// a hot path).  Anything else fails loudly: message on stderr and exit(1), LESGO's own convention for
// fatal errors (messages.f90:228-240) -- there is no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lesgo_gpu.h"
#include "fft_core.h"

using namespace lg;

namespace {

#ifdef LESGO_EMUL
LG_HD void sincospi_(double x, double* s, double* c) { *s = std::sin(3.14159265358979323846 * x); *c = std::cos(3.14159265358979323846 * x); }
#else
LG_D void sincospi_(double x, double* s, double* c) { sincospi(x, s, c); }
#endif

// One Stockham stage (radix R, Ns = product of the earlier radices) of `nbatch` transforms of length n.
// Element e of transform b sits at base + b*bstride + e*estride (in cplx units).  batch_fastest: consecutive
// threads take consecutive transforms (column transforms of a row-major plane) instead of consecutive
// butterflies.
template <int R>
__global__ void k_gfft_stage(const cplx* __restrict__ in, cplx* __restrict__ out, int n, int ns, long estride,
                             long bstride, int nbatch, int inverse, int batch_fastest) {
    const int T = n / R;
    const long total = long(T) * nbatch;
    for (long idx = long(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += long(gridDim.x) * blockDim.x) {
        int j, b;
        if (batch_fastest) { b = int(idx % nbatch); j = int(idx / nbatch); }
        else               { j = int(idx % T);      b = int(idx / T); }
        const cplx* pi = in + long(b) * bstride;
        cplx* po = out + long(b) * bstride;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = pi[long(j + r * T) * estride];
        const int k = j % ns;
        if (ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double s, c;
                sincospi_(-2.0 * double(long(k) * r) / double(long(ns) * R), &s, &c);
                const cplx w = make_double2(c, s);
                v[r] = inverse ? cmulc(v[r], w) : cmul(v[r], w);
            }
        }
        if (inverse) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
        }
        Dft<R>::run(v);
        if (inverse) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = cswap(v[r]);
        }
        const long j0 = long(j / ns) * ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) po[(j0 + long(r) * ns) * estride] = v[r];
    }
}

// real rows (row stride rs doubles) -> complex rows of length n0 (imaginary part 0)
__global__ void k_g_real_to_cplx(const double* __restrict__ in, cplx* __restrict__ out, int n0, int n1, long rs) {
    const long total = long(n0) * n1;
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const long y = i / n0, x = i % n0;
        out[i] = make_double2(in[y * rs + x], 0.0);
    }
}
// keep columns 0 .. n0/2 of complex rows of length n0
__global__ void k_g_take_half(const cplx* __restrict__ in, cplx* __restrict__ out, int n0, int n1) {
    const int lh = n0 / 2 + 1;
    const long total = long(lh) * n1;
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const long y = i / lh, x = i % lh;
        out[i] = in[y * n0 + x];
    }
}
// half spectrum rows (lh) -> full Hermitian rows (n0).  c2r semantics of FFTW's rdft2: the last (x) transform
// takes only the REAL parts of the k = 0 and k = n0/2 entries of each row.
__global__ void k_g_hermitian_rows(const cplx* __restrict__ in, cplx* __restrict__ out, int n0, int n1) {
    const int lh = n0 / 2 + 1;
    const long total = long(n0) * n1;
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const long y = i / n0;
        const int x = int(i % n0);
        cplx v;
        if (x < lh) {
            v = in[y * lh + x];
            if (x == 0 || 2 * x == n0) v.y = 0.0;
        } else {
            v = in[y * lh + (n0 - x)];
            v.y = -v.y;
        }
        out[i] = v;
    }
}
__global__ void k_g_cplx_to_real(const cplx* __restrict__ in, double* __restrict__ out, int n0, int n1, long rs) {
    const long total = long(n0) * n1;
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const long y = i / n0, x = i % n0;
        out[y * rs + x] = in[i].x;
    }
}

int grid_for(long n) {
    long b = (n + 255) / 256;
    return int(b < 1 ? 1 : (b > 148L * 32 ? 148L * 32 : b));
}

bool factor235(int n, std::vector<int>* radices) {
    if (n < 1) return false;
    while (n % 8 == 0) { radices->push_back(8); n /= 8; }
    while (n % 4 == 0) { radices->push_back(4); n /= 4; }
    while (n % 2 == 0) { radices->push_back(2); n /= 2; }
    while (n % 3 == 0) { radices->push_back(3); n /= 3; }
    while (n % 5 == 0) { radices->push_back(5); n /= 5; }
    return n == 1;
}

// all stages of `nbatch` transforms of length n; data starts in *a, result ends in *a (buffers swapped as needed)
void run_stages(cplx** a, cplx** b, int n, long estride, long bstride, int nbatch, bool inverse, bool batch_fastest) {
    std::vector<int> rad;
    factor235(n, &rad);
    int ns = 1;
    const int g = grid_for(long(n) * nbatch / 2);
    for (int R : rad) {
        const int inv = inverse ? 1 : 0, bf = batch_fastest ? 1 : 0;
        switch (R) {
            case 8: LG_LAUNCH(k_gfft_stage<8>, dim3(g), dim3(256), 0, nullptr, *a, *b, n, ns, estride, bstride, nbatch, inv, bf); break;
            case 5: LG_LAUNCH(k_gfft_stage<5>, dim3(g), dim3(256), 0, nullptr, *a, *b, n, ns, estride, bstride, nbatch, inv, bf); break;
            case 4: LG_LAUNCH(k_gfft_stage<4>, dim3(g), dim3(256), 0, nullptr, *a, *b, n, ns, estride, bstride, nbatch, inv, bf); break;
            case 3: LG_LAUNCH(k_gfft_stage<3>, dim3(g), dim3(256), 0, nullptr, *a, *b, n, ns, estride, bstride, nbatch, inv, bf); break;
            default: LG_LAUNCH(k_gfft_stage<2>, dim3(g), dim3(256), 0, nullptr, *a, *b, n, ns, estride, bstride, nbatch, inv, bf); break;
        }
        ns *= R;
        std::swap(*a, *b);
    }
}

struct FPlan {
    bool used = false;
    bool c2r = false;
    bool inplace = false;
    int n0 = 0, n1 = 0;
    int fast = -1;            // 0: ctx small grid, 1: ctx 3/2 grid, -1: generic
    lesgo_gpu_ctx* ctx = nullptr;
};
std::mutex g_mu;
std::vector<FPlan> g_plans(1);     // slot 0 stays unused: a zero handle is FFTW's "planning failed"
lesgo_gpu_ctx* g_bound = nullptr;
lesgo_gpu_dims g_bound_dims;
thread_local std::string g_shim_err;

int fail(const std::string& m) { g_shim_err = m; return 1; }

bool is_dev(const void* p) {
#ifdef LESGO_EMUL
    (void)p;
    return false;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
#endif
}

int generic_execute(const FPlan& p, double* in, double* out) {
    const int n0 = p.n0, n1 = p.n1, lh = n0 / 2 + 1;
    const long rs_real = p.inplace ? 2L * lh : n0;          // FFTW: in-place real rows are padded to 2*(n0/2+1)
    const size_t real_doubles = size_t(rs_real) * n1, cplx_doubles = size_t(2) * lh * n1;
    const size_t full = size_t(n0) * n1;
    cplx *A = nullptr, *B = nullptr;
    double* io = nullptr;
    const size_t io_doubles = real_doubles > cplx_doubles ? real_doubles : cplx_doubles;
    if (cudaMalloc(reinterpret_cast<void**>(&A), full * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&B), full * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&io), io_doubles * sizeof(double)) != cudaSuccess) {
        cudaFree(A); cudaFree(B); cudaFree(io);
        return fail("dfftw_execute: device allocation failed");
    }
    const cudaMemcpyKind k_in = is_dev(in) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind k_out = is_dev(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (!p.c2r) {
        cudaMemcpy(io, in, real_doubles * sizeof(double), k_in);
        LG_LAUNCH(k_g_real_to_cplx, dim3(grid_for(long(full))), dim3(256), 0, nullptr, io, A, n0, n1, rs_real);
        run_stages(&A, &B, n0, 1, n0, n1, false, false);                 // x: rows are contiguous
        LG_LAUNCH(k_g_take_half, dim3(grid_for(long(lh) * n1)), dim3(256), 0, nullptr, A, B, n0, n1);
        std::swap(A, B);
        run_stages(&A, &B, n1, lh, 1, lh, false, true);                  // y: columns of the (lh, n1) half spectrum
        cudaMemcpy(out, A, cplx_doubles * sizeof(double), k_out);
    } else {
        cudaMemcpy(A, in, cplx_doubles * sizeof(double), k_in);
        run_stages(&A, &B, n1, lh, 1, lh, true, true);
        LG_LAUNCH(k_g_hermitian_rows, dim3(grid_for(long(full))), dim3(256), 0, nullptr, A, B, n0, n1);
        std::swap(A, B);
        run_stages(&A, &B, n0, 1, n0, n1, true, false);
        LG_LAUNCH(k_g_cplx_to_real, dim3(grid_for(long(full))), dim3(256), 0, nullptr, A, io, n0, n1, rs_real);
        cudaMemcpy(out, io, real_doubles * sizeof(double), k_out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(A); cudaFree(B); cudaFree(io);
    if (e != cudaSuccess) return fail(std::string("dfftw_execute (generic path): ") + cudaGetErrorString(e));
    return 0;
}

[[noreturn]] void die(const char* where) {
    std::fprintf(stderr, "lesgo_gpu FFTW shim: %s: %s\n", where, g_shim_err.c_str());
    std::exit(1);
}

}  // namespace

extern "C" {

const char* lesgo_gpu_fftw_last_error(void) { return g_shim_err.c_str(); }

lesgo_gpu_ctx* lesgo_gpu_fftw_bound(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_bound;
}

int lesgo_gpu_fftw_bind(lesgo_gpu_ctx* ctx, const lesgo_gpu_dims* dims) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_bound = ctx;
    if (ctx) {
        if (!dims) return fail("lesgo_gpu_fftw_bind: dims required with a context");
        g_bound_dims = *dims;
    }
    return 0;
}

int lesgo_gpu_fftw_plan_2d(int c2r, int n0, int n1, int inplace, long long* plan) {
    if (!plan) return fail("null plan");
    *plan = 0;
    std::vector<int> r0, r1;
    if (n0 < 2 || n1 < 1 || !factor235(n0, &r0) || !factor235(n1, &r1))
        return fail("dfftw_plan_dft_*_2d: lengths " + std::to_string(n0) + " x " + std::to_string(n1) +
                    " do not factor into 2, 3, 5 (no generic transform for them, and no CPU fallback)");
    std::lock_guard<std::mutex> lk(g_mu);
    FPlan p;
    p.used = true; p.c2r = c2r != 0; p.inplace = inplace != 0; p.n0 = n0; p.n1 = n1; p.ctx = g_bound;
    if (g_bound && p.inplace) {
        if (n0 == g_bound_dims.nx && n1 == g_bound_dims.ny) p.fast = 0;
        else if (n0 == 3 * g_bound_dims.nx / 2 && n1 == 3 * g_bound_dims.ny / 2) p.fast = 1;
    }
    size_t slot = 0;
    for (size_t i = 1; i < g_plans.size(); ++i) if (!g_plans[i].used) { slot = i; break; }
    if (!slot) { g_plans.push_back(FPlan()); slot = g_plans.size() - 1; }
    g_plans[slot] = p;
    *plan = (long long)slot;
    return 0;
}

int lesgo_gpu_fftw_execute(long long plan, int c2r, double* in, double* out) {
    FPlan p;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (plan < 1 || size_t(plan) >= g_plans.size() || !g_plans[size_t(plan)].used)
            return fail("dfftw_execute: invalid or destroyed plan handle " + std::to_string(plan));
        p = g_plans[size_t(plan)];
    }
    if (p.c2r != (c2r != 0)) return fail("dfftw_execute: plan was made for the other direction");
    if (!in || !out) return fail("dfftw_execute: null array");
    if (p.inplace != (in == out))
        return fail("dfftw_execute: plan made in place must be executed in place and vice versa (FFTW new-array rule)");
    if (p.fast >= 0 && p.ctx && p.ctx == g_bound) {
        const int rc = p.c2r ? lesgo_gpu_fft_c2r(p.ctx, in, out, 1, p.fast) : lesgo_gpu_fft_r2c(p.ctx, in, out, 1, p.fast);
        if (rc) return fail(lesgo_gpu_last_error(p.ctx));
        return 0;
    }
    if (p.ctx && p.ctx == g_bound && g_bound_dims.device >= 0) cudaSetDevice(g_bound_dims.device);
    return generic_execute(p, in, out);
}

int lesgo_gpu_fftw_destroy(long long plan) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (plan < 1 || size_t(plan) >= g_plans.size() || !g_plans[size_t(plan)].used) return fail("dfftw_destroy_plan: invalid plan handle");
    g_plans[size_t(plan)] = FPlan();
    return 0;
}

// ---- the Fortran-mangled symbols ---------------------------------------------------------------------------------
// (a seventh argument, as fft.f90:114-121 passes, is ignored like FFTW's own wrapper ignores it)
void dfftw_plan_dft_r2c_2d_(long long* plan, const int* n0, const int* n1, double* in, double* out, const int* flags) {
    (void)flags;
    if (lesgo_gpu_fftw_plan_2d(0, *n0, *n1, in == out, plan)) die("dfftw_plan_dft_r2c_2d");
}
void dfftw_plan_dft_c2r_2d_(long long* plan, const int* n0, const int* n1, double* in, double* out, const int* flags) {
    (void)flags;
    if (lesgo_gpu_fftw_plan_2d(1, *n0, *n1, in == out, plan)) die("dfftw_plan_dft_c2r_2d");
}
void dfftw_execute_dft_r2c_(const long long* plan, double* in, double* out) {
    if (lesgo_gpu_fftw_execute(*plan, 0, in, out)) die("dfftw_execute_dft_r2c");
}
void dfftw_execute_dft_c2r_(const long long* plan, double* in, double* out) {
    if (lesgo_gpu_fftw_execute(*plan, 1, in, out)) die("dfftw_execute_dft_c2r");
}
void dfftw_destroy_plan_(long long* plan) {
    if (lesgo_gpu_fftw_destroy(*plan)) die("dfftw_destroy_plan");
    *plan = 0;
}

}  // extern "C"
