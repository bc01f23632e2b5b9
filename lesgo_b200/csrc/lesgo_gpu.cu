// lesgo_gpu.cu -- context, host-side orchestration and the extern "C" entry points of
// include/lesgo_gpu.h.  Mirrors, routine by routine, the reference's derivatives.f90,
// convec.f90, press_stag_array.f90, tridag_array.f90, fft.f90 and the time-loop glue of
// main.f90:155-344 / forcing.f90:149-244.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <unistd.h>
#include <vector>

#include "../../include/lesgo_gpu.h"
#include "comm.h"
#include "launch.h"
#include "prodfwd_kernels.h"
#include "lasd_kernels.h"
#include "turbine_kernels.h"
#include "tavg_kernels.h"

using namespace lg;

extern "C" lesgo_gpu_ctx* lesgo_gpu_fftw_bound(void);   // fftw_shim.cu (internal)

namespace {
const double kBogus = -1234567890.0;   // param.f90:93
thread_local std::string g_err;

}  // namespace

// developer experiment (LESGO_EXP_ALIAS=1): every plane of an INTERMEDIATE aliases plane 0, i.e. the
// intermediates are L2-resident by construction.  Results are garbage; the timings bound what
// keeping intermediates in L2 can buy.
static bool exp_alias() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_EXP_ALIAS"); v = (e && e[0] == '1') ? 1 : 0; }
    return v != 0;
}

struct lesgo_gpu_ctx {
    lesgo_gpu_dims d;
    int nx, ny, nz, lh, ld, nx2, ny2, lh_big, ld_big, nzt;
    long plane, plane_big, plane_bi;   // ld*ny, ld_big*ny2, ld*ny2 (big-y intermediate)
    bool bottom, top;
    int jzLo;
    double kxs, kys;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    long launches = 0;
    int device = 0;                    // CUDA device of this context (made current on every entry)
    int chunk = 0;                     // planes per pipelined chunk (0 = whole slab); see lesgo_gpu_create
    // twiddles: W_n[m] = exp(-2 pi i m/n); Wh_n[m] = exp(-2 pi i m/(2n)) for the real<->complex step
    cplx *Wx = nullptr, *Whx = nullptr, *Wxb = nullptr, *Whxb = nullptr, *Wy = nullptr, *Wyb = nullptr;
    // scratch
    double* sa[kMaxFields] = {nullptr};    // small spectra / intermediates, (ld, ny, 0:nz)
    double* sd[12] = {nullptr};            // lesgo_gpu_step's batched filt_da: 3 x spectra + 9 y-pass outputs
    double* bb[kMaxFields] = {nullptr};    // big-y intermediates, (ld, ny2, 0:nz)
    double* big[kMaxFields] = {nullptr};   // 3/2-grid physical fields, (ld_big, ny2, 0:nz)
    double* gam = nullptr;                 // tridiagonal gam(j) table (lh, ny, 0:nzt+1)
    double* pen_buf[2] = {nullptr};        // NCCL-path pencil buffers of a ragged ky split (ny % nproc != 0)
    double* work[13] = {nullptr};          // S11..S33, Nu_t, six stress-gradient temporaries (mode 1)
    double* lsq = nullptr;                 // l(k)**2 of the Smagorinsky length (nz+1)
    double* gtest = nullptr;               // test-filter kernel G_test (lh, ny)
    double* wplane[2] = {nullptr};         // filtered wall-adjacent u, v planes
    double* gtest2 = nullptr;              // second test-filter kernel G_test_test (sgs_model 5)
    int gcut[2] = {0, 0};                  // kx columns >= gcut[i] of kernel i are zero for every ky (sharp cut-off)
    double* lasd_buf[54] = {nullptr};      // lagrange_Sdep work fields (51) + 3 spectral scratch, lasd_chunk planes each
    double* lasd_tmp[4] = {nullptr};       // interpolag_Sdep's copies of F_LM, F_MM, F_QN, F_NN
    int lasd_chunk = 0;
    double* tavg_acc[TA_N] = {nullptr};    // running time averages (lesgo_gpu_tavg_compute)
    double* tavg_tmp[5] = {nullptr};       // w_uv, u_w, v_w, vortz, fza_uv
    double tavg_time = 0.0;
    // peer-memory pressure transposes (lesgo_gpu_comm_p2p_export / _import)
    double* p2p_buf = nullptr;             // two pencil buffers (alternating between solves), nproc blocks each
    size_t p2p_half = 0;                   // doubles per half
    double* p2p_pencil[8] = {nullptr};     // peers' pencil buffers in this rank's address space
    double* p2p_alt[8] = {nullptr};        // ... and their second halves
    double* p2p_flag = nullptr;            // device scalar for the stream-ordered barrier (NCCL all-reduce form)
    unsigned long long* p2p_sig[8] = {nullptr};   // peers' signal slots (16 per rank, after the two pencil halves)
    unsigned long long p2p_epoch = 0;      // barriers passed so far (the value the next barrier signals)
    bool p2p_flags_on = false;             // peer-memory flag barrier instead of the one-double NCCL all-reduce
    bool p2p_on = false;
    int p2p_parity = 0;
    std::vector<void*> ipc_opened;         // peers' buffers mapped with cudaIpcOpenMemHandle (closed in destroy)
    // actuator disks (lesgo_gpu_turbines_init)
    TurbSet turb;
    bool turb_on = false, turb_fz = false;
    int turb_adm = 0;
    std::vector<int> turb_nodes;           // prefix offsets of the disks' node lists (host copy of TurbSet::start)
    std::vector<void*> turb_allocs;        // the disks' device arrays: released when the disks are handed over again
    double cs_const = -1.0;                // value the whole Cs_opt2 field currently holds by assignment (< 0: unknown)
    double* turb_fzuv = nullptr;           // fza before interp_to_w_grid (uv nodes)
    int sgs_cfg = -1;                      // (sgs_model, ifilter) the tables above were built for
    double* fields[LG_NFIELDS] = {nullptr};
    std::vector<double*> staging;          // device staging for host-pointer arguments
    std::vector<size_t> staging_bytes;
    double* red_dev = nullptr;             // reductions
    double* red_host = nullptr;
    lg::Comm* comm = nullptr;
    std::vector<void*> allocs;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the host-pointer pipeline
    void* hp_ = nullptr;                            // Staged pipeline of the API call in progress (host arrays)
    // optional per-launch timing (lesgo_gpu_profile): CUDA events around every launch
    bool prof = false;
    struct ProfRec { const char* label; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;

    Lay lay() const { return Lay{plane, ld}; }
    Lay lay_big() const { return Lay{exp_alias() ? 0 : plane_big, ld_big}; }
    int fail(const std::string& m) { err = m; g_err = m; return 1; }
};

namespace {

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess)                                                        \
            return c->fail(std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

#ifdef LESGO_EMUL
struct ProfScope { ProfScope(lesgo_gpu_ctx*, const char*) {} };
#else
struct ProfScope {
    lesgo_gpu_ctx* c;
    size_t idx = 0;
    bool on;
    ProfScope(lesgo_gpu_ctx* c_, const char* label) : c(c_), on(c_->prof) {
        if (!on) return;
        lesgo_gpu_ctx::ProfRec r;
        r.label = label;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->stream);
        idx = c->prof_recs.size();
        c->prof_recs.push_back(r);
    }
    ~ProfScope() { if (on) cudaEventRecord(c->prof_recs[idx].b, c->stream); }
};
#endif

int dev_alloc(lesgo_gpu_ctx* c, double** p, size_t ndoubles) {
    if (*p) return 0;
    void* q = nullptr;
    CK(cudaMalloc(&q, ndoubles * sizeof(double)));
    CK(cudaMemsetAsync(q, 0, ndoubles * sizeof(double), c->stream));
    c->allocs.push_back(q);
    *p = static_cast<double*>(q);
    return 0;
}

int upload_table(lesgo_gpu_ctx* c, cplx** dst, const std::vector<cplx>& h) {
    void* q = nullptr;
    CK(cudaMalloc(&q, sizeof(cplx) * (h.size() + 1)));
    if (!h.empty()) CK(cudaMemcpy(q, h.data(), sizeof(cplx) * h.size(), cudaMemcpyHostToDevice));
    c->allocs.push_back(q);
    *dst = static_cast<cplx*>(q);
    return 0;
}

cplx unit_root(long m, long n) {   // exp(-2 pi i m / n), long double accuracy
    long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)(m % n) / (long double)n;
    return make_double2((double)cosl(ang), (double)sinl(ang));
}

// stage twiddles of Plan<n> in the layout fft_core.h's stage_load reads:
// for stage s >= 2 (radix R, Ns = product of earlier radices): T[(r-1)*Ns + k] = W_n^{k r n/(Ns R)}
int make_stage_twiddles(lesgo_gpu_ctx* c, cplx** dst, int n, bool columns = false) {
    PlanDesc d;
    if (!plan_lookup(n, &d)) return c->fail("no FFT plan for length " + std::to_string(n));
    std::vector<cplx> h;
    int ns = d.r[0];
    for (int s = 1; s < 4; ++s) {
        const int R = d.r[s];
        if (R == 1) break;
        const long step = n / (long(ns) * R);
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < ns; ++k) h.push_back(unit_root(long(k) * r * step, n));
        ns *= R;
    }
    if (int(h.size()) != d.twlen) return c->fail("internal: stage twiddle table length mismatch");
    int r1 = 0, r2 = 0;
    if (!columns && xplan2_lookup(n, &r1, &r2)) {
        // x passes: the table of the two-stage warp-scope inverse (wfft2_kernels.h) follows the Stockham stage tables:
        // rows r = 1..7, then r = 8, 16 of W_n^{j r}, j < 16
        for (int r = 1; r < 8; ++r)
            for (int j = 0; j < 16; ++j) h.push_back(unit_root(long(j) * r, n));
        for (int r = 8; r < r2; r += 8)
            for (int j = 0; j < 16; ++j) h.push_back(unit_root(long(j) * r, n));
    }
    if (columns && plan2_lookup(n, &r1, &r2)) {
        // y passes: the table of the two-stage column plan (fft_core.h fft_tile2) follows the Stockham stage tables:
        // rows r = 1..7, then r = 8, 16, 24 of W_n^{j r}, j < R1
        for (int r = 1; r < 8; ++r)
            for (int j = 0; j < r1; ++j) h.push_back(unit_root(long(j) * r, n));
        for (int r = 8; r < r2; r += 8)
            for (int j = 0; j < r1; ++j) h.push_back(unit_root(long(j) * r, n));
    }
    return upload_table(c, dst, h);
}

// W_{2m}^k, k = 0..m/2: the real<->half-complex untangling factors of a length-2m real row
int make_half_twiddles(lesgo_gpu_ctx* c, cplx** dst, int m) {
    std::vector<cplx> h(m / 2 + 1);
    for (int k = 0; k <= m / 2; ++k) h[k] = unit_root(k, 2L * m);
    return upload_table(c, dst, h);
}

int need_small(lesgo_gpu_ctx* c, int n) {
    for (int i = 0; i < n; ++i)
        if (dev_alloc(c, &c->sa[i], size_t(c->plane) * (c->nz + 1))) return 1;
    return 0;
}
int need_big(lesgo_gpu_ctx* c, int n) {
    for (int i = 0; i < n; ++i) {
        if (dev_alloc(c, &c->bb[i], size_t(c->plane_bi) * (c->nz + 1))) return 1;
        if (dev_alloc(c, &c->big[i], size_t(c->plane_big) * (c->nz + 1))) return 1;
    }
    return 0;
}

// ---- argument staging: host pointers are mirrored in device buffers --------------------
bool is_device_ptr(const void* p) {
#ifdef LESGO_EMUL
    (void)p;
    return true;
#else
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
#endif
}

// Host arrays are mirrored in device staging buffers.  Plain mode: whole-array H2D before, D2H after.
// Pipelined mode (field-shaped arrays of the big routines): inputs are uploaded in plane chunks on a
// copy stream, the compute stream waits only for the planes a pass needs (need()), and finished
// output planes are downloaded on a second copy stream while later planes are still being computed
// (done()), so H2D, compute and D2H overlap (PCIe is full duplex).
struct Staged {
    lesgo_gpu_ctx* c;
    struct Item {
        double* host; double* dev; size_t bytes; bool in, out; size_t skip;
        std::vector<char> sent;            // pipelined: planes already downloaded
    };
    std::vector<Item> items;
    size_t next_slot = 0;
    bool pipelined = false;
    int ch = 0;                            // planes per pipeline chunk
    int nplanes = 0;
    std::vector<cudaEvent_t> ev_in;        // ev_in[i]: input planes of chunk i are on the device
    std::vector<cudaEvent_t> ev_tmp;
    int waited = -1;                       // highest input chunk the compute stream already waits for
    explicit Staged(lesgo_gpu_ctx* c_, bool pipe = false) : c(c_) {
#ifndef LESGO_EMUL
        if (pipe && !std::getenv("LESGO_NO_PIPELINE")) {
            pipelined = true;
            nplanes = c->nz + 1;
            ch = nplanes > 64 ? 16 : (nplanes > 16 ? 8 : nplanes);
            if (!c->s_in) {
                cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
                cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
            }
        }
#else
        (void)pipe;
#endif
    }
    ~Staged() {
        for (auto e : ev_in) cudaEventDestroy(e);
        for (auto e : ev_tmp) cudaEventDestroy(e);
        if (c->hp_ == this) c->hp_ = nullptr;
    }
    // returns the device pointer to use for user pointer p holding `n` doubles; the first `skip`
    // doubles are never copied in either direction (arrays the reference declares 1:nz are passed
    // shifted down by one plane, so their "plane 0" is not the caller's memory)
    double* in(const double* p, size_t n, bool copy_in, bool copy_out, size_t skip = 0) {
        if (!p) return nullptr;
        if (is_device_ptr(p)) return const_cast<double*>(p);
        // same host array passed twice (in and out) shares one slot
        for (auto& it : items)
            if (it.host == p) { it.out = it.out || copy_out; return it.dev; }
        size_t slot = next_slot++;
        if (c->staging.size() <= slot) { c->staging.push_back(nullptr); c->staging_bytes.push_back(0); }
        if (c->staging_bytes[slot] < n * sizeof(double)) {
            if (c->staging[slot]) cudaFree(c->staging[slot]);
            void* q = nullptr;
            if (cudaMalloc(&q, n * sizeof(double)) != cudaSuccess) { c->fail("cudaMalloc staging"); return nullptr; }
            c->staging[slot] = static_cast<double*>(q);
            c->staging_bytes[slot] = n * sizeof(double);
        }
        double* d = c->staging[slot];
        if (!pipelined) {
            if (copy_in) cudaMemcpyAsync(d + skip, p + skip, (n - skip) * sizeof(double), cudaMemcpyHostToDevice, c->stream);
            else if (copy_out) cudaMemsetAsync(d, 0, n * sizeof(double), c->stream);   // pure output: defined pads, no H2D
        } else if (!copy_in && copy_out) {
            cudaMemsetAsync(d, 0, n * sizeof(double), c->stream);
        }
        Item it{const_cast<double*>(p), d, n * sizeof(double), copy_in, copy_out, skip, {}};
        if (pipelined) it.sent.assign(size_t(nplanes), 0);
        items.push_back(it);
        return d;
    }
    bool active() const { return pipelined && !items.empty(); }
    // pipelined: enqueue every input, chunk by chunk, on the upload stream
    void begin() {
        if (!active()) return;
        c->hp_ = this;
        const size_t pl = size_t(c->plane);
        const int nchunks = (nplanes + ch - 1) / ch;
        ev_in.resize(size_t(nchunks));
        for (int i = 0; i < nchunks; ++i) {
            const int ka = i * ch, kb = ka + ch < nplanes ? ka + ch : nplanes;
            for (auto& it : items) {
                if (!it.in) continue;
                size_t lo = size_t(ka) * pl, hi = size_t(kb) * pl;
                if (lo < it.skip) lo = it.skip;
                if (hi > lo) cudaMemcpyAsync(it.dev + lo, it.host + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, c->s_in);
            }
            cudaEventCreateWithFlags(&ev_in[size_t(i)], cudaEventDisableTiming);
            cudaEventRecord(ev_in[size_t(i)], c->s_in);
        }
    }
    // the next launches on the compute stream read input planes <= kmax
    void need(int kmax) {
        if (!active() || ev_in.empty()) return;
        if (kmax >= nplanes) kmax = nplanes - 1;
        const int i = kmax / ch;
        if (i <= waited) return;
        cudaStreamWaitEvent(c->stream, ev_in[size_t(i)], 0);   // chunks are uploaded in order
        waited = i;
    }
    void need_all() { need(nplanes - 1); }
    // planes [pa, pb) of output array `dev` are final: download them behind the compute stream
    void done(const double* dev, int pa, int pb) {
        if (!active() || pb <= pa) return;
        for (auto& it : items) {
            if (it.dev != dev || !it.out) continue;
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            cudaEventRecord(e, c->stream);
            cudaStreamWaitEvent(c->s_out, e, 0);
            ev_tmp.push_back(e);
            const size_t pl = size_t(c->plane);
            size_t lo = size_t(pa) * pl, hi = size_t(pb) * pl;
            if (lo < it.skip) lo = it.skip;
            if (hi > lo) cudaMemcpyAsync(it.host + lo, it.dev + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c->s_out);
            for (int k = pa; k < pb; ++k) it.sent[size_t(k)] = 1;
        }
    }
    // planes [pa, pb) of `dev` were modified again (BOGUS fills): they must be (re)sent at the end
    void touch(const double* dev, int pa, int pb) {
        if (!active()) return;
        for (auto& it : items)
            if (it.dev == dev && it.out)
                for (int k = pa; k < pb && k < nplanes; ++k) it.sent[size_t(k)] = 0;
    }
    int finish() {
        if (active()) {
            // everything not downloaded yet goes out behind the compute stream
            cudaStreamSynchronize(c->s_out);
            const size_t pl = size_t(c->plane);
            for (auto& it : items) {
                if (!it.out) continue;
                int k = 0;
                while (k < nplanes) {
                    if (it.sent[size_t(k)]) { ++k; continue; }
                    int k1 = k;
                    while (k1 < nplanes && !it.sent[size_t(k1)]) ++k1;
                    size_t lo = size_t(k) * pl, hi = size_t(k1) * pl;
                    if (lo < it.skip) lo = it.skip;
                    if (hi > lo) cudaMemcpyAsync(it.host + lo, it.dev + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
                    k = k1;
                }
            }
            cudaStreamSynchronize(c->s_in);
            cudaError_t e = cudaStreamSynchronize(c->stream);
            c->hp_ = nullptr;
            if (e != cudaSuccess) return c->fail(std::string("stream sync: ") + cudaGetErrorString(e));
            e = cudaGetLastError();
            if (e != cudaSuccess) return c->fail(std::string("kernel: ") + cudaGetErrorString(e));
            return 0;
        }
        bool any = false;
        for (auto& it : items)
            if (it.out) { cudaMemcpyAsync(it.host + it.skip, it.dev + it.skip, it.bytes - it.skip * sizeof(double), cudaMemcpyDeviceToHost, c->stream); any = true; }
        if (any || !items.empty()) {
            cudaError_t e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return c->fail(std::string("stream sync: ") + cudaGetErrorString(e));
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return c->fail(std::string("kernel: ") + cudaGetErrorString(e));
        return 0;
    }
};

Staged* HP(const lesgo_gpu_ctx* c) {
    Staged* h = static_cast<Staged*>(c->hp_);
    return (h && h->active()) ? h : nullptr;
}

// ---- pass wrappers ------------------------------------------------------------------------
int grid1d(long n) {
    long b = (n + kBlock - 1) / kBlock;
    long cap = 148L * 16;
    return int(b < 1 ? 1 : (b > cap ? cap : b));
}

template <class Pro>
int xfwd(lesgo_gpu_ctx* c, bool bigx, const Pro& pro, int nf, double* const* dst, long dplane, int drow,
         int ncol, int nyrows, int k0, int k1, int write_nyq = 0) {
    XfOut o;
    for (int i = 0; i < nf; ++i) o.dst[i] = dst[i];
    o.plane = dplane; o.row = drow; o.ncol = ncol; o.write_nyq = write_nyq; o.ring = 0;
    if (exp_alias()) o.ring = 1;
    if (k1 <= k0) return 0;
    ProfScope ps_(c, bigx ? "xfwd_big" : "xfwd");
    int rc = launch_xfwd<Pro>(bigx ? c->nx2 : c->nx, pro, nf, o, nyrows, k0, k1 - k0,
                              bigx ? c->Wxb : c->Wx, bigx ? c->Whxb : c->Whx, c->stream);
    if (rc) return c->fail("unsupported nx for x-forward pass");
    c->launches++;
    return 0;
}

// optional time-stepping glue fused into an x-inverse epilogue (EpiStore modes 1 and 2)
struct Fuse {
    int mode = 0;
    const double* divt[3] = {nullptr, nullptr, nullptr};
    double* rhs_f[3] = {nullptr, nullptr, nullptr};
    double* u[3] = {nullptr, nullptr, nullptr};
    double force[3] = {0, 0, 0};
    int kmax[3] = {0, 0, 0};
    int first_step = 0;
    double dt = 0, t1 = 0, t2 = 0;
    const double* fa[3] = {nullptr, nullptr, nullptr};   // applied body force (actuator disks)
    int kfa = 0;
    // dpdz fusion (press): RHSz, w and the projection range
    double* rhsz = nullptr;
    double* w = nullptr;
    int kproj = 0, kproj_end = 0;
};

int xinv(lesgo_gpu_ctx* c, bool bigx, const double* const* src, long splane, int srow, int ncol, int nf,
         double* const* dst, const Lay& dl, int nyrows, int k0, int k1, int pad = 1, const Fuse* fz = nullptr) {
    XiSrc in;
    for (int i = 0; i < nf; ++i) in.src[i] = src[i];
    in.plane = splane; in.row = srow; in.ncol = ncol; in.ring = 0;
    if (exp_alias()) in.ring = 1;
    if (k1 <= k0) return 0;
    ProfScope ps_(c, bigx ? "xinv_big" : "xinv");
    int rc;
    if (fz && fz->mode && !bigx) {
        EpiFused epi;
        std::memset(&epi, 0, sizeof(epi));
        for (int i = 0; i < nf; ++i) epi.dst[i] = dst[i];
        epi.lay = dl; epi.nx = c->nx; epi.pad = pad;
        epi.mode = fz->mode; epi.first_step = fz->first_step; epi.dt = fz->dt; epi.t1 = fz->t1; epi.t2 = fz->t2;
        for (int i = 0; i < 3; ++i) { epi.divt[i] = fz->divt[i]; epi.rhs_f[i] = fz->rhs_f[i]; epi.u[i] = fz->u[i]; epi.force[i] = fz->force[i]; epi.kmax[i] = fz->kmax[i]; epi.fa[i] = fz->fa[i]; }
        epi.kfa = fz->kfa;
        rc = launch_xinv_fused(c->nx, in, epi, nf, nyrows, k0, k1 - k0, c->Wx, c->Whx, c->stream);
    } else {
        EpiStore epi;
        for (int i = 0; i < nf; ++i) epi.dst[i] = dst[i];
        epi.lay = dl; epi.nx = bigx ? c->nx2 : c->nx; epi.pad = pad;
        rc = launch_xinv(bigx ? c->nx2 : c->nx, in, epi, nf, nyrows, k0, k1 - k0,
                         bigx ? c->Wxb : c->Wx, bigx ? c->Whxb : c->Whx, c->stream);
    }
    if (rc) return c->fail("unsupported nx for x-inverse pass");
    c->launches++;
    return 0;
}

YArgs yargs(lesgo_gpu_ctx* c, long splane, int srow, long dplane, int drow, int ncols, int k0) {
    YArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nout = 1;
    a.src_plane = splane; a.src_row = srow; a.dst_plane = dplane; a.dst_row = drow;
    a.ncols = ncols; a.k0 = k0;
    a.kxs = c->kxs; a.kys = c->kys;
    a.table = nullptr; a.table2 = nullptr; a.table_row = 0;
    a.zero_col = -1; a.keep_nyq_row = 0;
    return a;
}

int ypass(lesgo_gpu_ctx* c, int nin, int nout, const YArgs& a, int nf, int k0, int k1) {
    if (k1 <= k0) return 0;
    ProfScope ps_(c, nin == nout ? "ypass_deriv" : (nout == 0 ? "ypass_fwd" : (nin == 0 ? "ypass_inv" : (nin < nout ? "ypass_pad" : "ypass_trunc"))));
    auto W = [&](int n) -> const cplx* { return n == c->ny ? c->Wy : (n == c->ny2 ? c->Wyb : nullptr); };
    if (exp_alias()) { YArgs b = a; b.src_ring = b.dst_ring = 1; int rc = launch_ypass(nin, nout, b, nf, k1 - k0, nin ? W(nin) : W(nout), nout ? W(nout) : W(nin), c->stream); c->launches++; return rc; }
    int rc = launch_ypass(nin, nout, a, nf, k1 - k0, nin ? W(nin) : W(nout), nout ? W(nout) : W(nin), c->stream);
    if (rc) return c->fail("unsupported ny for y pass");
    c->launches++;
    return 0;
}

int glue_fused(lesgo_gpu_ctx* c, int mode, double* rhs, const double* b, double* rhs_f, double* u, int k0, int k1,
               int kproj, int first_step, double force, double dt, double t1, double t2) {
    if (k1 <= k0) return 0;
    ProfScope ps_(c, "glue");
    LG_LAUNCH(k_glue_fused, dim3(grid1d(long(c->lh) * c->ny * (k1 - k0))), dim3(kBlock), 0, c->stream, mode, rhs, b, rhs_f,
              u, c->lay(), c->nx, c->ny, k0, k1, kproj, first_step, force, dt, t1, t2);
    c->launches++;
    return 0;
}

// Consecutive fills inside a FillGroup scope are issued as ONE launch when the scope ends.
struct FillGroup {
    lesgo_gpu_ctx* c;
    FillList l;
    FillGroup* prev;
    static thread_local FillGroup* cur;
    explicit FillGroup(lesgo_gpu_ctx* c_) : c(c_), prev(cur) { l.count = 0; cur = this; }
    void flush() {
        if (l.count == 0) return;
        long nmax = 0;
        for (int i = 0; i < l.count; ++i) nmax = l.n[i] > nmax ? l.n[i] : nmax;
        ProfScope ps_(c, "fill");
        LG_LAUNCH(k_fill_multi, dim3(grid1d(nmax), l.count), dim3(kBlock), 0, c->stream, l);
        c->launches++;
        l.count = 0;
    }
    void add(double* p, long n, double v) {
        if (l.count == FillList::kMax) flush();
        l.p[l.count] = p; l.n[l.count] = n; l.v[l.count] = v; ++l.count;
    }
    ~FillGroup() { flush(); cur = prev; }
};
thread_local FillGroup* FillGroup::cur = nullptr;

int fill(lesgo_gpu_ctx* c, double* f, long plane, int k0, int k1, double v) {
    if (k1 <= k0) return 0;
    if (Staged* h = HP(c)) h->touch(f, k0, k1);
    if (FillGroup::cur && FillGroup::cur->c == c) {
        FillGroup::cur->add(f + long(k0) * plane, plane * (k1 - k0), v);
        return 0;
    }
    ProfScope ps_(c, "fill");
    LG_LAUNCH(k_fill, dim3(grid1d(plane * (k1 - k0))), dim3(kBlock), 0, c->stream, f, plane, k0, k1, v);
    c->launches++;
    return 0;
}

// ---- derivatives.f90 --------------------------------------------------------------------------
// which: bit 0 = f itself (filt_da), bit 1 = d/dx, bit 2 = d/dy
int chunk_of(const lesgo_gpu_ctx* c, int divisor) {
    if (Staged* h = HP(c)) return h->ch;            // host-array pipeline: its chunking rules
    if (c->chunk <= 0) return c->nz + 1;
    int ch = c->chunk / divisor;
    return ch < 1 ? 1 : ch;
}

// bigy != nullptr (filt_da inside lesgo_gpu_step): the y pass also writes the 3/2-rule padded inverse
// transform of the field's spectrum -- (ld, 3ny/2, 0:nz), what convec's own pad pass would produce
// from the filtered field -- which convec then takes over instead of transforming u, v, w again
int spectral_deriv(lesgo_gpu_ctx* c, const double* f, double* fout, double* dfdx, double* dfdy, double* bigy = nullptr) {
    if (need_small(c, 4)) return 1;
    const int nz = c->nz;
    ProScale pro;
    pro.src[0] = f; pro.lay = c->lay(); pro.scale = 1.0 / (double(c->nx) * double(c->ny));
    double* d0[1] = {c->sa[0]};
    YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, c->nx / 2, 0);
    a.fld[0].src = c->sa[0];
    const double* xs[3];
    double* xd[3];
    int n = 0;
    auto mid = [&](int i) { return c->sa[1 + i]; };
    if (fout) { a.fld[0].out[n] = YOutSpec{mid(n), Y_COPY}; xs[n] = mid(n); xd[n] = fout; ++n; }
    if (dfdx) { a.fld[0].out[n] = YOutSpec{mid(n), Y_IKX}; xs[n] = mid(n); xd[n] = dfdx; ++n; }
    if (dfdy) { a.fld[0].out[n] = YOutSpec{mid(n), Y_IKY}; xs[n] = mid(n); xd[n] = dfdy; ++n; }
    a.nout = n;
    // plane chunks: the x->y->x passes of a chunk run back to back so the two spectral
    // intermediates are still in L2 when the next pass reads them
    const int ch = chunk_of(c, 1);
    for (int ka = 0; ka < nz + 1; ka += ch) {
        const int kb = ka + ch < nz + 1 ? ka + ch : nz + 1;
        Staged* hp = HP(c);
        if (hp) hp->need(kb - 1);
        if (xfwd(c, false, pro, 1, d0, c->plane, c->ld, c->nx / 2, c->ny, ka, kb)) return 1;
        a.k0 = ka;
        if (bigy) {
            a.fld[0].out2 = bigy; a.dst2_plane = c->plane_bi; a.dst2_row = c->ld;
            ProfScope ps_(c, "ypass_deriv");
            if (launch_ypass_pad2(c->ny, a, 1, kb - ka, c->Wy, c->Wyb, c->stream)) return c->fail("unsupported ny for y pass");
            c->launches++;
        } else if (ypass(c, c->ny, c->ny, a, 1, ka, kb)) return 1;
        if (xinv(c, false, xs, c->plane, c->ld, c->nx / 2, n, xd, c->lay(), c->ny, ka, kb)) return 1;
        if (hp) for (int i = 0; i < n; ++i) hp->done(xd[i], ka, kb);
    }
    return 0;
}

bool batch_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_BATCH"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// main.f90:161-163 in one go: filt_da of u, v and w as ONE x-forward launch (3 fields), ONE y pass (3 fields, 3 + 1
// outputs each) and ONE x-inverse launch (9 fields) instead of three of each.  Same kernels, same arithmetic; the
// point is the tail of the persistent grids and the launch count, which matter most on the short slabs of a
// many-GPU run (profiles/r4_experiments.md).
int spectral_deriv_uvw(lesgo_gpu_ctx* c, double* const* F, bool bigy) {
    const int nz = c->nz;
    for (int i = 0; i < 12; ++i)
        if (dev_alloc(c, &c->sd[i], size_t(c->plane) * (nz + 1))) return 1;
    ProScale pro;
    pro.src[0] = F[LG_U]; pro.src[1] = F[LG_V]; pro.src[2] = F[LG_W];
    pro.lay = c->lay(); pro.scale = 1.0 / (double(c->nx) * double(c->ny));
    if (xfwd(c, false, pro, 3, c->sd, c->plane, c->ld, c->nx / 2, c->ny, 0, nz + 1)) return 1;
    YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, c->nx / 2, 0);
    for (int f = 0; f < 3; ++f) {
        a.fld[f].src = c->sd[f];
        a.fld[f].out[0] = YOutSpec{c->sd[3 + 3 * f], Y_COPY};
        a.fld[f].out[1] = YOutSpec{c->sd[4 + 3 * f], Y_IKX};
        a.fld[f].out[2] = YOutSpec{c->sd[5 + 3 * f], Y_IKY};
        if (bigy) a.fld[f].out2 = c->bb[f];
    }
    a.nout = 3;
    if (bigy) {
        a.dst2_plane = c->plane_bi; a.dst2_row = c->ld;
        ProfScope ps_(c, "ypass_deriv");
        if (launch_ypass_pad2(c->ny, a, 3, nz + 1, c->Wy, c->Wyb, c->stream)) return c->fail("unsupported ny for y pass");
        c->launches++;
    } else if (ypass(c, c->ny, c->ny, a, 3, 0, nz + 1)) return 1;
    const double* xs[9];
    for (int i = 0; i < 9; ++i) xs[i] = c->sd[3 + i];
    double* xd[9] = {F[LG_U], F[LG_DUDX], F[LG_DUDY], F[LG_V], F[LG_DVDX], F[LG_DVDY], F[LG_W], F[LG_DWDX], F[LG_DWDY]};
    return xinv(c, false, xs, c->plane, c->ld, c->nx / 2, 9, xd, c->lay(), c->ny, 0, nz + 1);
}

int ddz_uv(lesgo_gpu_ctx* c, const double* f, double* dfdz) {
    const int nz = c->nz;
    Staged* hp = HP(c);
    const int ch = hp ? hp->ch : nz + 1;
    for (int ka = 1; ka < nz + 1; ka += ch) {                       // dfdz(k) = (f(k) - f(k-1))/dz, k = 1..nz
        const int kb = ka + ch < nz + 1 ? ka + ch : nz + 1;
        if (hp) hp->need(kb - 1);
        {
            ProfScope ps_(c, "ddz");
            LG_LAUNCH(k_ddz, dim3(grid1d(long(c->nx / 2) * c->ny * (kb - ka))), dim3(kBlock), 0, c->stream, f, dfdz, c->lay(),
                      c->nx, c->ny, ka, kb, -1, 0, 1.0 / c->d.dz);
            c->launches++;
        }
        if (hp) hp->done(dfdz, ka, kb);
    }
    FillGroup fg_(c);
    fill(c, dfdz, c->plane, 0, 1, kBogus);                           // derivatives.f90:236-238
    if (c->bottom) fill(c, dfdz, c->plane, 1, 2, kBogus);            // :255-257
    if (c->top) fill(c, dfdz, c->plane, nz, nz + 1, kBogus);         // :258-260
    return 0;
}

int ddz_w(lesgo_gpu_ctx* c, const double* f, double* dfdz) {
    const int nz = c->nz;
    Staged* hp = HP(c);
    const int ch = hp ? hp->ch : nz + 1;
    for (int ka = 0; ka < nz; ka += ch) {                           // dfdz(k) = (f(k+1) - f(k))/dz, k = 0..nz-1
        const int kb = ka + ch < nz ? ka + ch : nz;
        if (hp) hp->need(kb);
        {
            ProfScope ps_(c, "ddz");
            LG_LAUNCH(k_ddz, dim3(grid1d(long(c->nx / 2) * c->ny * (kb - ka))), dim3(kBlock), 0, c->stream, f, dfdz, c->lay(),
                      c->nx, c->ny, ka, kb, 0, 1, 1.0 / c->d.dz);
            c->launches++;
        }
        if (hp) hp->done(dfdz, ka, kb);
    }
    FillGroup fg_(c);
    if (c->bottom) fill(c, dfdz, c->plane, 0, 1, kBogus);            // :303-305
    fill(c, dfdz, c->plane, nz, nz + 1, kBogus);                     // :308
    return 0;
}

bool reuse_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_REUSE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
bool prodfwd_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_PRODFWD"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
// planes per z chunk of the z-marching product pass: long enough that the one-plane overlap between
// chunks is cheap, short enough that ny2 * nchunks work items balance over the persistent blocks
int prod_chunk(int nz) {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("LESGO_PROD_CHUNK"); v = e ? std::atoi(e) : 0; }
    if (v > 0) return v;
    const int n = nz - 1;
    if (n <= 40) return n < 1 ? 1 : n;
    const int nch = (n + 63) / 64;
    return (n + nch - 1) / nch;
}

// ---- convec.f90 --------------------------------------------------------------------------------
int convec(lesgo_gpu_ctx* c, const double* u, const double* v, const double* w, const double* dudy,
           const double* dudz, const double* dvdx, const double* dvdz, const double* dwdx,
           const double* dwdy, double* RHSx, double* RHSy, double* RHSz, const Fuse* fz = nullptr,
           bool uvw_ready = false) {
    if (need_small(c, 6)) return 1;
    if (need_big(c, 6)) return 1;
    const int nz = c->nz, nxh = c->nx / 2;
    const double cs = 1.0 / (double(c->nx) * double(c->ny));
    ProScale ps;
    ps.src[0] = u; ps.src[1] = v; ps.src[2] = w; ps.lay = c->lay(); ps.scale = cs;
    ProVort pv;
    pv.dudy = dudy; pv.dudz = dudz; pv.dvdx = dvdx; pv.dvdz = dvdz; pv.dwdx = dwdx; pv.dwdy = dwdy;
    pv.lay = c->lay(); pv.scale = cs; pv.nz = nz; pv.bottom = c->bottom; pv.top = c->top;
    pv.lbc_mom = c->d.lbc_mom; pv.ubc_mom = c->d.ubc_mom;
    ProConvec pc;
    pc.u = c->big[0]; pc.v = c->big[1]; pc.w = c->big[2]; pc.o1 = c->big[3]; pc.o2 = c->big[4]; pc.o3 = c->big[5];
    pc.lay = c->lay_big(); pc.scale = 1.0 / (double(c->nx2) * double(c->ny2));
    pc.nz = nz; pc.bottom = c->bottom; pc.top = c->top; pc.jzLo = c->jzLo;
    double* out[3] = {RHSx, RHSy, RHSz};
    // Plane chunks, software-pipelined by one plane: the products of plane p need the 3/2-grid
    // fields at p-1, p, p+1, so chunk [ka,kb) first brings planes [ka,kb) to the 3/2 grid and then
    // forms the products of planes [ka-1, kb-1) -- everything a chunk touches is still in L2.
    const int ch = chunk_of(c, 4);
    for (int ka = 0; ka < nz + 1; ka += ch) {
        const int kb = ka + ch < nz + 1 ? ka + ch : nz + 1;
        const int va = ka < 1 ? 1 : ka;                 // vorticity exists on planes 1..nz
        if (Staged* hp = HP(c)) hp->need(kb);           // the wall-plane vorticity reads one plane up
        if (uvw_ready) {
            // Spectral reuse (lesgo_gpu_step only): filt_da's y pass already wrote the 3/2-grid y transforms
            // of the filtered u, v, w into bb[0..2] (k_ypass<ny, ny, multi, 3ny/2>), i.e. what steps (1)-(2)
            // would recompute from the filtered fields (same spectrum, same padding, same inverse
            // transform).  Only the vorticity goes through (1)-(2).
            // (Forming the vorticity's x spectra from kept derivative spectra and z differences was
            // measured too: the extra strided loads in the y pass cost more than the x-forward pass
            // they replace, profiles/r2_experiments.md.)
            if (xfwd(c, false, pv, 3, c->sa + 3, c->plane, c->ld, nxh, c->ny, va, kb)) return 1;
            YArgs b = yargs(c, c->plane, c->ld, c->plane_bi, c->ld, nxh, va);
            for (int i = 0; i < 3; ++i) { b.fld[i].src = c->sa[3 + i]; b.fld[i].out[0] = YOutSpec{c->bb[3 + i], Y_COPY}; }
            if (ypass(c, c->ny, c->ny2, b, 3, va, kb)) return 1;
        } else {
        // (1) u, v, w and the vorticity to half spectra                          convec.f90:73-82, 97-158
        if (xfwd(c, false, ps, 3, c->sa, c->plane, c->ld, nxh, c->ny, ka, kb)) return 1;
        if (xfwd(c, false, pv, 3, c->sa + 3, c->plane, c->ld, nxh, c->ny, va, kb)) return 1;
        // (2) y forward, padd (fft.f90:43-71), y inverse on the 3/2 grid
        {
            YArgs a = yargs(c, c->plane, c->ld, c->plane_bi, c->ld, nxh, ka);
            for (int i = 0; i < 3; ++i) { a.fld[i].src = c->sa[i]; a.fld[i].out[0] = YOutSpec{c->bb[i], Y_COPY}; }
            if (ypass(c, c->ny, c->ny2, a, 3, ka, kb)) return 1;
            YArgs b = yargs(c, c->plane, c->ld, c->plane_bi, c->ld, nxh, va);
            for (int i = 0; i < 3; ++i) { b.fld[i].src = c->sa[3 + i]; b.fld[i].out[0] = YOutSpec{c->bb[3 + i], Y_COPY}; }
            if (ypass(c, c->ny, c->ny2, b, 3, va, kb)) return 1;
        }
        }
        // (3) x inverse on the 3/2 grid: only kx < nx/2 carries data               :90-92, 165-167
        if (xinv(c, true, c->bb, c->plane_bi, c->ld, nxh, 3, c->big, c->lay_big(), c->ny2, ka, kb)) return 1;
        if (xinv(c, true, c->bb + 3, c->plane_bi, c->ld, nxh, 3, c->big + 3, c->lay_big(), c->ny2, va, kb)) return 1;
        // products of the planes whose upper neighbour is now available
        const int pa = ka - 1 < 1 ? 1 : ka - 1;
        const int pb = kb == nz + 1 ? nz + 1 : kb - 1;
        if (pb <= pa) continue;
        // (4) products fused into the x forward pass on the 3/2 grid               :172-305
        if (prodfwd_enabled() && !HP(c) && c->chunk <= 0 && nz >= 2) {
            // one kernel marching up z per 3/2-grid row: every operand read once (prodfwd_kernels.h)
            ProdArgs b;
            for (int i = 0; i < 6; ++i) b.src[i] = c->big[i];
            for (int i = 0; i < 3; ++i) b.dst[i] = c->bb[i];
            const Lay lb = c->lay_big();
            b.splane = lb.plane; b.srow = lb.row; b.dplane = c->plane_bi; b.drow = c->ld;
            b.ny2 = c->ny2; b.nz = nz; b.bottom = c->bottom; b.top = c->top; b.jzLo = c->jzLo;
            b.chunk = prod_chunk(nz); b.nchunks = (nz - 1 + b.chunk - 1) / b.chunk;
            b.scale = 1.0 / (double(c->nx2) * double(c->ny2));
            {
                ProfScope ps_(c, "xfwd_big");
                if (launch_prodfwd(c->nx2, b, c->Wxb, c->Whxb, c->stream)) return c->fail("unsupported nx for the 3/2-grid product pass");
                c->launches++;
            }
            { FillGroup fg_(c); for (int i = 0; i < 3; ++i) fill(c, c->bb[i], c->plane_bi, nz, nz + 1, 0.0); }   // cc(nz) = 0, :262-268
        } else
        if (xfwd(c, true, pc, 3, c->bb, c->plane_bi, c->ld, nxh, c->ny2, pa, pb)) return 1;
        // (5) y forward on the 3/2 grid, unpadd (fft.f90:74-99), y inverse          :206-213
        {
            YArgs a = yargs(c, c->plane_bi, c->ld, c->plane, c->ld, nxh, pa);
            for (int i = 0; i < 3; ++i) { a.fld[i].src = c->bb[i]; a.fld[i].out[0] = YOutSpec{c->sa[i], Y_COPY}; }
            if (ypass(c, c->ny2, c->ny, a, 3, pa, pb)) return 1;
        }
        // (6) x inverse -> RHS
        if (xinv(c, false, c->sa, c->plane, c->ld, nxh, 3, out, c->lay(), c->ny, pa, pb, 1, fz)) return 1;
        if (Staged* hp = HP(c)) for (int i = 0; i < 3; ++i) hp->done(out[i], pa, pb);
    }
    // :319-332
    FillGroup fg_(c);
    fill(c, RHSx, c->plane, 0, 1, kBogus); fill(c, RHSy, c->plane, 0, 1, kBogus); fill(c, RHSz, c->plane, 0, 1, kBogus);
    fill(c, RHSx, c->plane, nz, nz + 1, kBogus); fill(c, RHSy, c->plane, nz, nz + 1, kBogus);
    if (!c->top) fill(c, RHSz, c->plane, nz, nz + 1, kBogus);
    return 0;
}

// ---- press_stag_array.f90 ------------------------------------------------------------------------
// one-off (the gam table is built once per context): did any mode hit a zero pivot?
int pivot_check(lesgo_gpu_ctx* c, const double* flag) {
    int h = 0;
    CK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h) { c->gam = nullptr; return c->fail("tridag_array failed: zero pivot (tridag_array.f90:55-60,101-108)"); }
    return 0;
}

int tridag_setup(lesgo_gpu_ctx* c, TriGeom& g) {
    g.lh = c->lh; g.ny = c->ny; g.nzt = c->nzt; g.row = c->ld; g.plane = c->plane;
    g.gplane = long(c->lh) * c->ny; g.kxs = c->kxs; g.kys = c->kys; g.dz = c->d.dz;
    if (!c->gam) {
        if (dev_alloc(c, &c->gam, size_t(g.gplane) * (c->nzt + 2))) return 1;
        const int nm = (c->lh - 1) * c->ny;
        double* flag = nullptr;
        if (dev_alloc(c, &flag, 1)) return 1;          // zeroed
        LG_LAUNCH(k_tridag_setup, dim3((nm + 127) / 128), dim3(128), 0, c->stream, g, c->gam, reinterpret_cast<int*>(flag));
        c->launches++;
        if (pivot_check(c, flag)) return 1;
    }
    return 0;
}

int press(lesgo_gpu_ctx* c, const double* u, const double* v, const double* w, const double* divtz, double dt,
          double tadv1, double* p, double* dpdx, double* dpdy, double* dpdz, const Fuse* fz = nullptr) {
    if (c->d.nproc > 1 && !c->comm) return c->fail("press_stag_array: nproc > 1 needs lesgo_gpu_comm_init first");
    if (need_small(c, 6)) return 1;
    const int nz = c->nz, nxh = c->nx / 2;
    const double cst = 1.0 / (double(c->nx) * double(c->ny));
    const double cst2 = cst / tadv1 / dt;                          // press_stag_array.f90:51-52
    // (1) x forward of u*/(tadv1 dt): planes 1..nz-1, + w(nz) on the top rank      :77-103
    // (2) y forward in place, Nyquist row zeroed                                   :129-146
    ProScale ps;
    ps.src[0] = u; ps.src[1] = v; ps.src[2] = w; ps.lay = c->lay(); ps.scale = cst2;
    const int ch = chunk_of(c, 2);
    for (int ka = 1; ka < nz; ka += ch) {
        const int kb = ka + ch < nz ? ka + ch : nz;
        if (Staged* hp = HP(c)) hp->need(kb - 1);
        if (xfwd(c, false, ps, 3, c->sa, c->plane, c->ld, nxh, c->ny, ka, kb)) return 1;
        YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, nxh, ka);
        for (int i = 0; i < 3; ++i) { a.fld[i].src = c->sa[i]; a.fld[i].out[0] = YOutSpec{c->sa[i], Y_COPY}; }
        if (ypass(c, c->ny, 0, a, 3, ka, kb)) return 1;
    }
    // boundary planes: w(nz) on the top rank, divtz at the walls                    :100-126
    if (Staged* hp = HP(c)) hp->need_all();
    ProScale pd; pd.src[0] = divtz; pd.lay = c->lay(); pd.scale = cst;
    double* d3[1] = {c->sa[3]};
    if (c->top) {
        ProScale pw; pw.src[0] = w; pw.lay = c->lay(); pw.scale = cst2;
        double* d[1] = {c->sa[2]};
        if (xfwd(c, false, pw, 1, d, c->plane, c->ld, nxh, c->ny, nz, nz + 1)) return 1;
        if (xfwd(c, false, pd, 1, d3, c->plane, c->ld, nxh, c->ny, nz, nz + 1)) return 1;
        YArgs b = yargs(c, c->plane, c->ld, c->plane, c->ld, nxh, nz);
        b.fld[0].src = c->sa[2]; b.fld[0].out[0] = YOutSpec{c->sa[2], Y_COPY};
        b.fld[1].src = c->sa[3]; b.fld[1].out[0] = YOutSpec{c->sa[3], Y_COPY};
        if (ypass(c, c->ny, 0, b, 2, nz, nz + 1)) return 1;
    }
    if (c->bottom) {
        if (xfwd(c, false, pd, 1, d3, c->plane, c->ld, nxh, c->ny, 1, 2)) return 1;
        YArgs e = yargs(c, c->plane, c->ld, c->plane, c->ld, nxh, 1);
        e.fld[0].src = c->sa[3]; e.fld[0].out[0] = YOutSpec{c->sa[3], Y_COPY};
        if (ypass(c, c->ny, 0, e, 1, 1, 2)) return 1;
    }
    // (3) tridiagonal solve + k=0 chain                                           :149-239
    double* phat = c->sa[4];
    if (c->d.nproc == 1) {
        TriGeom g;
        if (tridag_setup(c, g)) return 1;
        const int nm = (c->lh - 1) * c->ny;
        ProfScope ps_(c, "tridag");
        LG_LAUNCH(k_tridag_fused, dim3((nm + 127) / 128), dim3(128), 0, c->stream, g, c->sa[0], c->sa[1],
                  c->sa[2], c->sa[3] + c->plane, c->sa[3] + c->plane * nz, c->gam, c->sa[4]);
        c->launches++;
    } else {
        // slabs -> pencils -> Thomas -> slabs (see PencilGeom in ops.h); replaces the rank-serial
        // pipeline of tridag_array.f90:22-162 and the halo/chain messages C4-C7.
        PencilGeom g;
        g.lh = c->lh; g.ny = c->ny; g.ld = c->ld; g.nz = nz; g.nproc = c->d.nproc; g.coord = c->d.coord;
        g.cy = (c->ny + c->d.nproc - 1) / c->d.nproc; g.plane = c->plane; g.kxs = c->kxs; g.kys = c->kys; g.dz = c->d.dz;
        g.p2p = c->p2p_on ? 1 : 0;
        // the two halves of every rank's buffer alternate from call to call: a rank may already push the next
        // solve's rows into a peer while a third rank is still pulling the previous result out of it
        const int par = c->p2p_parity;
        if (c->p2p_on) c->p2p_parity ^= 1;
        for (int q = 0; q < 8; ++q) g.pencil[q] = par ? c->p2p_alt[q] : c->p2p_pencil[q];
        // stream-ordered barrier of the peer-memory path: every rank's pushes are complete (kernel boundary)
        // before any rank's next kernel reads them
        auto barrier = [&]() -> int {
            ProfScope ps_(c, "p2p_barrier");
            if (c->p2p_flags_on) {
                // flag barrier in peer memory: signal every rank's slot [coord] with the new epoch, wait for all of ours
                P2PSig sg;
                for (int q = 0; q < 8; ++q) sg.peer[q] = c->p2p_sig[q];
                sg.rank = c->d.coord; sg.nproc = c->d.nproc; sg.epoch = ++c->p2p_epoch;
                LG_LAUNCH(k_p2p_barrier, dim3(1), dim3(32), 0, c->stream, sg);
                c->launches++;
                return 0;
            }
            if (c->comm->allreduce_sum_dev(c->p2p_flag, 1, c->stream)) return c->fail(c->comm->error());
            return 0;
        };
        // NCCL path: send / receive buffers.  The small spectra serve when ny divides evenly; a ragged
        // split (cy * nproc > ny) pads every block to cy rows and may need more than a field's worth.
        double *nsend = c->sa[4], *nrecv = c->sa[5];
        if (!c->p2p_on && size_t(g.block()) * g.nproc > size_t(c->plane) * (nz + 1)) {
            for (int i = 0; i < 2; ++i)
                if (dev_alloc(c, &c->pen_buf[i], size_t(g.block()) * g.nproc)) return 1;
            nsend = c->pen_buf[0]; nrecv = c->pen_buf[1];
        }
        double* pencil = c->p2p_on ? c->p2p_buf + (par ? c->p2p_half : 0) : nrecv;
        double* ret = nsend;                                      // NCCL path only
        {   // rH_z(1) of coord+1 -> rH_z(nz) of coord                             :184-185
            const double* sb[1] = {c->sa[2] + c->plane};
            double* rb[1] = {c->sa[2] + c->plane * nz};
            int dest[1] = {c->d.coord - 1}, src[1] = {c->d.coord + 1};
            size_t cnt[1] = {size_t(c->plane)};
            ProfScope ps_(c, "halo");
            if (c->comm->exchange(1, sb, dest, rb, src, cnt, c->stream)) return c->fail(c->comm->error());
        }
        if (!c->gam) {
            if (dev_alloc(c, &c->gam, size_t(c->nzt + 2) * g.cy * c->lh)) return 1;
            const int nm = (c->lh - 1) * g.cy;
            double* flag = nullptr;
            if (dev_alloc(c, &flag, 1)) return 1;      // zeroed
            LG_LAUNCH(k_tridag_setup_pencil, dim3((nm + 127) / 128), dim3(128), 0, c->stream, g, c->nzt, c->gam, reinterpret_cast<int*>(flag));
            c->launches++;
            if (pivot_check(c, flag)) return 1;
        }
        {
            ProfScope ps_(c, "press_pack");
            LG_LAUNCH(k_press_pack, dim3(grid1d(long(c->lh - 1) * c->ny * nz)), dim3(kBlock), 0, c->stream, g, c->sa[0],
                      c->sa[1], c->sa[2], c->sa[3] + c->plane, c->sa[3] + c->plane * nz, nsend);
            c->launches++;
        }
        if (c->p2p_on) { if (barrier()) return 1; }
        else {
            ProfScope ps_(c, "alltoall");
            if (c->comm->alltoall(nsend, nrecv, size_t(g.block()), c->stream)) return c->fail(c->comm->error());
        }
        {
            const int nm = (c->lh - 1) * g.cy;
            ProfScope ps_(c, "tridag");
            // the sweep is local and in place on both paths
            // LESGO_PENCIL_PIPE=1: request the operands of the next batch of rows before processing the one in hand.
            // Measured SLOWER on 8 B200s (0.318 against 0.240 ms): the sweep is bound by its dependent instruction
            // chain (address arithmetic + an FP64 division per row) on ~7 warps per SM, not by the loads, and the
            // second register set only lengthens it (profiles/r4_experiments.md).  Off by default.
            static const bool pipe = std::getenv("LESGO_PENCIL_PIPE") && std::getenv("LESGO_PENCIL_PIPE")[0] == '1';
            if (pipe) LG_LAUNCH(k_tridag_pencil<true>, dim3((2 * nm + 127) / 128), dim3(128), 0, c->stream, g, c->nzt, c->gam, pencil);
            else LG_LAUNCH(k_tridag_pencil<false>, dim3((2 * nm + 127) / 128), dim3(128), 0, c->stream, g, c->nzt, c->gam, pencil);
            c->launches++;
        }
        if (c->p2p_on) { if (barrier()) return 1; }
        else {
            ProfScope ps_(c, "alltoall");
            if (c->comm->alltoall(nrecv, nsend, size_t(g.block()), c->stream)) return c->fail(c->comm->error());
        }
        {
            ProfScope ps_(c, "press_unpack");
            LG_LAUNCH(k_press_unpack, dim3(grid1d(long(c->lh) * c->ny * nz)), dim3(kBlock), 0, c->stream, g, ret, c->sa[3]);
            c->launches++;
        }
        phat = c->sa[3];
        {   // p(nz-1) of coord -> p(0) of coord+1                                  :241-246
            const double* sb[1] = {phat + c->plane * (nz - 1)};
            double* rb[1] = {phat};
            int dest[1] = {c->d.coord + 1}, src[1] = {c->d.coord - 1};
            size_t cnt[1] = {size_t(c->plane)};
            ProfScope ps_(c, "halo");
            if (c->comm->exchange(1, sb, dest, rb, src, cnt, c->stream)) return c->fail(c->comm->error());
        }
    }
    // (4) y inverse of p, i kx p, i ky p (oddballs dropped), x inverse              :248-273
    // (5) dpdz = (p(k) - p(k-1))/dz                                                 :276-288
    {
        YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, nxh, 0);
        a.fld[0].src = phat;
        a.fld[0].out[0] = YOutSpec{c->sa[0], Y_COPY};
        a.fld[0].out[1] = YOutSpec{c->sa[1], Y_IKX};
        a.fld[0].out[2] = YOutSpec{c->sa[2], Y_IKY};
        a.nout = 3;
        const double* s0[1] = {c->sa[0]};
        double* o0[1] = {p};
        const double* s1[2] = {c->sa[1], c->sa[2]};
        double* o1[2] = {dpdx, dpdy};
        const int pend = c->top ? nz + 1 : nz;          // p is transformed on planes 0..pend-1
        for (int ka = 0; ka < nz + 1; ka += ch) {
            const int kb = ka + ch < nz + 1 ? ka + ch : nz + 1;
            a.k0 = ka;
            if (ypass(c, 0, c->ny, a, 1, ka, kb)) return 1;
            Staged* hp = HP(c);
            if (xinv(c, false, s0, c->plane, c->ld, nxh, 1, o0, c->lay(), c->ny, ka, kb < pend ? kb : pend)) return 1;
            if (hp) hp->done(p, ka, kb < pend ? kb : pend);
            const int da = ka < 1 ? 1 : ka, db = kb < nz ? kb : nz;
            if (xinv(c, false, s1, c->plane, c->ld, nxh, 2, o1, c->lay(), c->ny, da, db, 1, fz)) return 1;
            if (hp) { hp->done(dpdx, da, db); hp->done(dpdy, da, db); }
            const int za = ka < 1 ? 1 : ka, zb = kb < pend ? kb : pend;
            if (zb > za) {
                ProfScope ps_(c, "dpdz");
                LG_LAUNCH(k_dpdz, dim3(grid1d(long(nxh) * c->ny * (zb - za))), dim3(kBlock), 0, c->stream, p, dpdz, c->lay(),
                          c->nx, c->ny, za, zb, c->d.dz, fz ? fz->rhsz : nullptr, fz ? fz->w : nullptr,
                          fz ? fz->kproj : 0, fz ? fz->kproj_end : 0, fz ? fz->dt : 0.0, fz ? fz->t1 : 0.0);
                c->launches++;
                if (hp) hp->done(dpdz, za, zb);
            }
        }
    }
    FillGroup fg_(c);
    fill(c, dpdx, c->plane, nz, nz + 1, kBogus);
    fill(c, dpdy, c->plane, nz, nz + 1, kBogus);
    if (!c->top) { fill(c, p, c->plane, nz, nz + 1, kBogus); fill(c, dpdz, c->plane, nz, nz + 1, kBogus); }
    return 0;
}

double* field(lesgo_gpu_ctx* c, int id) {
    if (id < 0 || id >= LG_NFIELDS) return nullptr;
    if (!c->fields[id]) dev_alloc(c, &c->fields[id], size_t(c->plane) * (c->nz + 1));
    return c->fields[id];
}

// ---- wallstress + sgs_stag (constant coefficient) + divstress on the device ---------
int build_sgs_tables(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp) {
    const int cfg = sp->sgs_model * 16 + sp->ifilter;
    if (c->sgs_cfg == cfg && c->lsq) return 0;
    const int nz = c->nz;
    const double dx = c->d.L_x / c->nx, dy = c->d.L_y / c->ny, dz = c->d.dz;
    const double delta = std::pow(dx * dy * dz, 1.0 / 3.0);                  // sgs_param.f90:187
    // l(k), sgs_stag_util.f90:87-179 (sgs_model 1) / :183 (l = delta otherwise)
    std::vector<double> l(nz + 1, delta);
    const int lb = c->d.lbc_mom, ub = c->d.ubc_mom;
    if (sp->sgs_model == 1 && !(lb == 0 && ub == 0)) {
        const double Co = sp->Co, n = sp->wall_damp_exp, vonk = sp->vonk;
        auto damp = [&](double zz) { return std::pow(std::pow(Co, n) * std::pow(vonk * zz, -n) + std::pow(delta, -n), -1.0 / n); };
        int jmin = 1, jmax = nz;
        if (lb > 0 && c->bottom) { l[1] = damp(0.5 * dz); jmin = 2; }
        if (ub > 0 && c->top) { l[nz] = damp(0.5 * dz); jmax = nz - 1; }
        for (int jz = jmin; jz <= jmax; ++jz) {
            double zz;
            if (lb > 0 && ub == 0) zz = ((jz - 1) + c->d.coord * (nz - 1)) * dz;
            else if (lb > 0 && ub > 0) { zz = ((jz - 1) + c->d.coord * (nz - 1)) * dz; zz = std::fmin(zz, (nz - 1) * c->d.nproc * dz - zz); }
            else zz = ((c->d.nproc - c->d.coord) * (nz - 1) - (jz - 1)) * dz;
            l[jz] = damp(zz);
        }
    }
    for (auto& v : l) v = v * v;
    if (!c->lsq && dev_alloc(c, &c->lsq, nz + 1)) return 1;
    CK(cudaMemcpyAsync(c->lsq, l.data(), sizeof(double) * (nz + 1), cudaMemcpyHostToDevice, c->stream));
    // G_test (alpha_test = 2) and, for sgs_model 5, G_test_test (alpha_test_test = 4), test_filtermodule.f90:38-123
    std::vector<double> G(size_t(c->lh) * c->ny);
    const double pi = 3.14159265358979323846;
    for (int which = 0; which < (sp->sgs_model == 5 ? 2 : 1); ++which) {
    const double dt_ = (which == 0 ? 2.0 : 4.0) * std::sqrt(dx * dy), kc2 = (pi / dt_) * (pi / dt_);
    for (int jy = 0; jy < c->ny; ++jy)
        for (int jx = 0; jx < c->lh; ++jx) {
            double kx = c->kxs * jx, ky = c->kys * double(jy < c->ny / 2 ? jy : jy - c->ny);
            if (jx == c->lh - 1 || jy == c->ny / 2) { kx = 0.0; ky = 0.0; }
            const double k2 = kx * kx + ky * ky;
            double g = 1.0 / (double(c->nx) * double(c->ny));
            if (sp->ifilter == 1) { if (k2 >= kc2) g = 0.0; }
            else if (sp->ifilter == 2) g = std::exp(-(dt_ * dt_) * k2 / (4.0 * 6.0)) * g;
            else if (sp->ifilter == 3) g = (std::sin(kx * dt_ / 2.0) * std::sin(ky * dt_ / 2.0) + 1e-8) / (kx * dt_ / 2.0 * ky * dt_ / 2.0 + 1e-8) * g;
            if (jx == c->lh - 1 || jy == c->ny / 2) g = 0.0;
            G[size_t(jy) * c->lh + jx] = g;
        }
    // columns of the kernel that are zero for every ky: the filtered spectrum is zero there, so the passes
    // neither compute nor move them (spectral cut-off: 3/4 resp. 7/8 of the half spectrum)
    int cut = 0;
    for (int jy = 0; jy < c->ny; ++jy)
        for (int jx = cut; jx < c->lh; ++jx)
            if (G[size_t(jy) * c->lh + jx] != 0.0) cut = jx + 1;
    c->gcut[which] = cut < 1 ? 1 : (cut > c->nx / 2 ? c->nx / 2 : cut);
    double** gdst = which == 0 ? &c->gtest : &c->gtest2;
    if (!*gdst && dev_alloc(c, gdst, G.size())) return 1;
    CK(cudaMemcpyAsync(*gdst, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    }
    c->sgs_cfg = cfg;
    return 0;
}

// LESGO_FILTER_PRUNE=0: transform and move the kernels' all-zero kx columns too (A/B and test switch)
static bool prune_enabled() {
    static const bool on = [] { const char* e = std::getenv("LESGO_FILTER_PRUNE"); return !e || std::atoi(e) != 0; }();
    return on;
}

// test_filter of ONE plane (test_filtermodule.f90:126-146): src plane -> dst plane
int filter_plane(lesgo_gpu_ctx* c, const double* src, double* dst) {
    if (need_small(c, 2)) return 1;
    ProScale ps; ps.src[0] = src; ps.lay = c->lay(); ps.scale = 1.0;
    double* d0[1] = {c->sa[0]};
    const int nc = prune_enabled() ? c->gcut[0] : c->nx / 2;
    if (xfwd(c, false, ps, 1, d0, c->plane, c->ld, nc, c->ny, 0, 1)) return 1;
    YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, nc, 0);
    a.fld[0].src = c->sa[0]; a.fld[0].out[0] = YOutSpec{c->sa[1], Y_TABLE};
    a.table = c->gtest; a.table_row = c->lh;
    if (ypass(c, c->ny, c->ny, a, 1, 0, 1)) return 1;
    const double* s0[1] = {c->sa[1]};
    double* o0[1] = {dst};
    return xinv(c, false, s0, c->plane, c->ld, nc, 1, o0, c->lay(), c->ny, 0, 1);
}

int wallstress(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp, double* const* F, bool with_tau) {
    // wallstress.f90:47-255; with_tau = false (core mode) only sets dudz, dvdz on the wall planes
    const int nz = c->nz;
    const double h = 0.5 * c->d.dz;
    const int g1 = grid1d(long(c->nx) * c->ny);
    for (int side = 0; side < 2; ++side) {
        const bool bot = side == 0;
        if (bot ? !c->bottom : !c->top) continue;
        const int bc = bot ? c->d.lbc_mom : c->d.ubc_mom;
        const int ksrc = bot ? 1 : nz - 1, kdst = bot ? 1 : nz;
        if (bc == 0) {
            fill(c, F[LG_DUDZ], c->plane, kdst, kdst + 1, 0.0); fill(c, F[LG_DVDZ], c->plane, kdst, kdst + 1, 0.0);
            if (with_tau) { fill(c, F[LG_TXZ], c->plane, kdst, kdst + 1, 0.0); fill(c, F[LG_TYZ], c->plane, kdst, kdst + 1, 0.0); }
        } else if (bc == 1) {
            ProfScope ps_(c, "wall");
            LG_LAUNCH(k_wall_dns, dim3(grid1d(long(c->nx / 2) * c->ny)), dim3(kBlock), 0, c->stream, F[LG_U], F[LG_V],
                      F[LG_DUDZ], F[LG_DVDZ], c->lay(), c->nx, c->ny, ksrc, kdst, bot ? sp->ubot : sp->utop, bot ? 1.0 : -1.0, h);
            c->launches++;
            if (with_tau) {
                LG_LAUNCH(k_wall_tau_dns, dim3(g1), dim3(kBlock), 0, c->stream, F[LG_DUDZ], F[LG_DVDZ], F[LG_TXZ], F[LG_TYZ],
                          c->lay(), c->nx, c->ny, kdst, sp->nu_molec_nd);
                c->launches++;
            }
        } else if (bc == 2) {
            if (!with_tau) return c->fail("lesgo_gpu_step: equilibrium wall model needs mode 1");
            for (int i = 0; i < 2; ++i)
                if (dev_alloc(c, &c->wplane[i], size_t(c->plane))) return 1;
            if (filter_plane(c, F[LG_U] + c->plane * ksrc, c->wplane[0])) return 1;
            if (filter_plane(c, F[LG_V] + c->plane * ksrc, c->wplane[1])) return 1;
            ProfScope ps_(c, "wall");
            LG_LAUNCH(k_wall_equil, dim3(g1), dim3(kBlock), 0, c->stream, F[LG_U], F[LG_V], c->wplane[0], c->wplane[1],
                      F[LG_DUDZ], F[LG_DVDZ], F[LG_TXZ], F[LG_TYZ], c->lay(), c->nx, c->ny, ksrc, kdst, bot ? 1.0 : -1.0,
                      sp->vonk, std::log(h / sp->zo), h * sp->vonk);
            c->launches++;
        } else {
            return c->fail("lesgo_gpu_step: lbc_mom/ubc_mom = 3 (integral wall model) is out of scope");
        }
    }
    return 0;
}

int plane_exchange(lesgo_gpu_ctx* c, const double* send, int dest, double* recv, int src) {
    if (!c->comm) return 0;
    const double* sb[1] = {send};
    double* rb[1] = {recv};
    int d[1] = {dest}, s[1] = {src};
    size_t cnt[1] = {size_t(c->plane)};
    ProfScope ps_(c, "halo");
    if (c->comm->exchange(1, sb, d, rb, s, cnt, c->stream)) return c->fail(c->comm->error());
    return 0;
}

int sum3(lesgo_gpu_ctx* c, double* out, const double* a, const double* b, const double* cc, int k0, int k1, int zero_pad) {
    if (k1 <= k0) return 0;
    ProfScope ps_(c, "glue");
    LG_LAUNCH(k_sum3, dim3(grid1d(long(c->lh) * c->ny * (k1 - k0))), dim3(kBlock), 0, c->stream, out, a, b, cc, c->lay(),
              c->nx, c->ny, k0, k1, zero_pad);
    c->launches++;
    return 0;
}

// ---- Lagrangian scale-dependent dynamic model -------------------------------------------
// test_filter AND test_test_filter (test_filtermodule.f90:126-168) of nf <= 3 fields on planes k0..k1-1: the
// forward x transform is shared by the two filters.  Arrays are addressed by absolute plane index.
// `sp2`: nf more spectral scratch arrays (addressed by absolute plane index like everything else here).
int filter_fields(lesgo_gpu_ctx* c, int nf, const double* const* src, double* const* dst1, double* const* dst2,
                  double* const* sp2, int k0, int k1) {
    if (need_small(c, 6)) return 1;
    ProScale ps;
    for (int i = 0; i < nf; ++i) ps.src[i] = src[i];
    ps.lay = c->lay(); ps.scale = 1.0;
    double* xs[3] = {c->sa[0], c->sa[1], c->sa[2]};
    // kx columns where a kernel vanishes for every ky are neither transformed nor moved: the x transform writes
    // nc1 columns, the y pass inverts nc1 (G_test) and nc2 (G_test_test) of them, the x inverses read that many
    const bool prune = prune_enabled();
    const int nc1 = prune ? (c->gcut[0] > c->gcut[1] ? c->gcut[0] : c->gcut[1]) : c->nx / 2;
    const int nc2 = prune ? c->gcut[1] : c->nx / 2;
    if (xfwd(c, false, ps, nf, xs, c->plane, c->ld, nc1, c->ny, k0, k1)) return 1;
    // one y pass: forward transform once, G_test and G_test_test applied to the same spectrum, two inverses
    YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, nc1, k0);
    a.nout = 2; a.ncols2 = nc2;
    for (int i = 0; i < nf; ++i) {
        a.fld[i].src = c->sa[i];
        a.fld[i].out[0] = YOutSpec{c->sa[3 + i], Y_TABLE};
        a.fld[i].out[1] = YOutSpec{sp2[i], Y_TABLE2};
    }
    a.table = c->gtest; a.table2 = c->gtest2; a.table_row = c->lh;
    if (ypass(c, c->ny, c->ny, a, nf, k0, k1)) return 1;
    const double* s0[6];
    double* d0[6];
    for (int i = 0; i < nf; ++i) { s0[i] = c->sa[3 + i]; d0[i] = dst1[i]; s0[nf + i] = sp2[i]; d0[nf + i] = dst2[i]; }
    if (nc1 == nc2) return xinv(c, false, s0, c->plane, c->ld, nc1, 2 * nf, d0, c->lay(), c->ny, k0, k1);
    if (xinv(c, false, s0, c->plane, c->ld, nc1, nf, d0, c->lay(), c->ny, k0, k1)) return 1;
    return xinv(c, false, s0 + nf, c->plane, c->ld, nc2, nf, d0 + nf, c->lay(), c->ny, k0, k1);
}

// lagrange_Sdep (lagrange_Sdep.f90:22-430) including interpolag_Sdep (interpolag_Sdep.f90:21-268); Sij in c->work[0..5]
int lagrange_sdep(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp, double* const* F) {
    const int nz = c->nz;
    const size_t nfield = size_t(c->plane) * (nz + 1);
    LasdGeom g;
    g.nx = c->nx; g.ny = c->ny; g.nz = nz; g.coord = c->d.coord; g.nproc = c->d.nproc;
    g.bottom = c->bottom; g.top = c->top; g.lbc_mom = c->d.lbc_mom; g.ubc_mom = c->d.ubc_mom;
    g.dx = c->d.L_x / c->nx; g.dy = c->d.L_y / c->ny; g.dz = c->d.dz;
    g.L_x = c->d.L_x; g.L_y = c->d.L_y; g.L_z = (c->d.nz_tot - 1) * c->d.dz;
    double* FL[4] = {F[LG_F_LM], F[LG_F_MM], F[LG_F_QN], F[LG_F_NN]};
    auto sync4 = [&]() -> int {                                   // mpi_sync_real_array(F_*, 0, MPI_SYNC_DOWNUP)
        if (!c->comm) return 0;
        ProfScope ps_(c, "halo");
        for (int f = 0; f < 4; ++f)
            if (c->comm->sync_planes(FL[f], c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
        return 0;
    };
    // interpolag_Sdep.f90:69-72 copies, :84-241 backward trajectories (k = nz on the top rank only), :244-249 sync
    {
        LasdInterpArgs ia;
        ia.u = F[LG_U]; ia.v = F[LG_V]; ia.w = F[LG_W]; ia.lagran_dt = sp->lagran_dt;
        for (int f = 0; f < 4; ++f) {
            if (dev_alloc(c, &c->lasd_tmp[f], nfield)) return 1;
            CK(cudaMemcpyAsync(c->lasd_tmp[f], FL[f], nfield * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            ia.T[f] = c->lasd_tmp[f]; ia.F[f] = FL[f];
        }
        const int k1 = c->top ? nz + 1 : nz;
        ProfScope ps_(c, "lasd_interp");
        LG_LAUNCH(k_interpolag, dim3(grid1d(long(c->nx) * c->ny * (k1 - 1))), dim3(kBlock), 0, c->stream, ia, g, c->lay(), 1, k1);
        c->launches++;
    }
    if (sync4()) return 1;
    // the per-plane part (:83-413), `chunk` planes at a time so the 51 work fields stay small
    if (!c->lasd_chunk) {
        const char* e = std::getenv("LESGO_LASD_CHUNK");
        int ch = e ? std::atoi(e) : 32;
        c->lasd_chunk = ch < 1 ? 1 : (ch > nz ? nz : ch);
        for (int i = 0; i < 54; ++i)
            if (dev_alloc(c, &c->lasd_buf[i], size_t(c->plane) * c->lasd_chunk)) return 1;
    }
    const double dx = g.dx, dy = g.dy;
    const double delta = std::pow(dx * dy * g.dz, 1.0 / 3.0);     // sgs_param.f90:187
    for (int k0 = 1; k0 <= nz; k0 += c->lasd_chunk) {
        const int k1 = (k0 + c->lasd_chunk < nz + 1) ? k0 + c->lasd_chunk : nz + 1;
        // work fields addressed by the absolute plane index
        double* B[54];
        for (int i = 0; i < 54; ++i) B[i] = c->lasd_buf[i] - long(k0) * c->plane;
        double** SP2 = B + 51;
        double **A = B, **Tb = B + 9, **Th = B + 18, **Sb = B + 27, **Sh = B + 33, **SSb = B + 39, **SSh = B + 45;
        const int g1 = grid1d(long(c->nx) * c->ny * (k1 - k0));
        {
            LasdPrepArgs pa;
            pa.u = F[LG_U]; pa.v = F[LG_V]; pa.w = F[LG_W];
            for (int i = 0; i < 9; ++i) pa.A[i] = A[i];
            ProfScope ps_(c, "lasd_prep");
            LG_LAUNCH(k_lasd_prep, dim3(g1), dim3(kBlock), 0, c->stream, pa, g, c->lay(), k0, k1);
            c->launches++;
        }
        for (int i = 0; i < 9; i += 3)
            if (filter_fields(c, 3, A + i, Tb + i, Th + i, SP2, k0, k1)) return 1;
        for (int i = 0; i < 6; i += 3)
            if (filter_fields(c, 3, c->work + i, Sb + i, Sh + i, SP2, k0, k1)) return 1;
        {
            LasdSSArgs sa;
            for (int i = 0; i < 6; ++i) { sa.S[i] = c->work[i]; sa.SS[i] = A[i]; }
            ProfScope ps_(c, "lasd_ss");
            LG_LAUNCH(k_lasd_ss, dim3(g1), dim3(kBlock), 0, c->stream, sa, c->lay(), c->nx, c->ny, k0, k1);
            c->launches++;
        }
        for (int i = 0; i < 6; i += 3)
            if (filter_fields(c, 3, A + i, SSb + i, SSh + i, SP2, k0, k1)) return 1;
        {
            LasdFinalArgs fa;
            for (int i = 0; i < 9; ++i) { fa.Tb[i] = Tb[i]; fa.Th[i] = Th[i]; }
            for (int i = 0; i < 6; ++i) { fa.Sb[i] = Sb[i]; fa.Sh[i] = Sh[i]; fa.SSb[i] = SSb[i]; fa.SSh[i] = SSh[i]; }
            fa.F_LM = FL[0]; fa.F_MM = FL[1]; fa.F_QN = FL[2]; fa.F_NN = FL[3]; fa.Cs = F[LG_CS_OPT2];
            fa.delta = delta; fa.lagran_dt = sp->lagran_dt;
            fa.beta_exp = std::log(2.0) / (std::log(4.0) - std::log(2.0));       // log(tf1) / (log(tf2) - log(tf1))
            fa.init_F = sp->lasd_init_F ? 1 : 0;
            ProfScope ps_(c, "lasd_final");
            LG_LAUNCH(k_lasd_final, dim3(grid1d(long(c->ld) * c->ny * (k1 - k0))), dim3(kBlock), 0, c->stream, fa, g, c->lay(), k0, k1);
            c->launches++;
        }
    }
    return sync4();                                               // :417-420
}

int sgs_and_divstress(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp, double* const* F) {
    const int nz = c->nz, coord = c->d.coord;
    for (int i = 0; i < 13; ++i)
        if (dev_alloc(c, &c->work[i], size_t(c->plane) * (nz + 1))) return 1;
    if (build_sgs_tables(c, sp)) return 1;
    SgsParams p;
    p.nz = nz; p.bottom = c->bottom; p.top = c->top; p.lbc_mom = c->d.lbc_mom; p.ubc_mom = c->d.ubc_mom;
    p.sgs = c->d.sgs; p.nu = sp->nu_molec_nd; p.Cs_opt2 = sp->sgs_model == 1 ? sp->Co * sp->Co : 0.03;
    // calc_Sij needs dwdz(nz) = dwdz(1) of the rank above (sgs_stag_util.f90:611-614)
    if (plane_exchange(c, F[LG_DWDZ] + c->plane, coord - 1, F[LG_DWDZ] + c->plane * nz, coord + 1)) return 1;
    SijArgs sa;
    sa.dudx = F[LG_DUDX]; sa.dudy = F[LG_DUDY]; sa.dudz = F[LG_DUDZ]; sa.dvdx = F[LG_DVDX]; sa.dvdy = F[LG_DVDY];
    sa.dvdz = F[LG_DVDZ]; sa.dwdx = F[LG_DWDX]; sa.dwdy = F[LG_DWDY]; sa.dwdz = F[LG_DWDZ];
    for (int i = 0; i < 6; ++i) sa.S[i] = c->work[i];
    sa.Nu_t = c->work[6]; sa.lsq = c->lsq;
    {
        ProfScope ps_(c, "sgs");
        LG_LAUNCH(k_sij_nut, dim3(grid1d(long(c->nx) * c->ny * nz)), dim3(kBlock), 0, c->stream, sa, p, c->lay(), c->nx, c->ny, 1, nz + 1);
        c->launches++;
    }
    if (c->d.sgs && sp->sgs_model == 1) {
        // sgs_stag_util.f90:94 assigns the whole ARRAY Cs_opt2 = Co**2 every step: tavg%compute (time_average.f90:258)
        // and the restart file (io.f90:1204-1211) see that field.  The value is constant: written when it changes.
        const double cs = sp->Co * sp->Co;
        if (c->cs_const != cs) {
            double* f = field(c, LG_CS_OPT2);
            if (!f || fill(c, f, c->plane, 0, nz + 1, cs)) return 1;
            c->cs_const = cs;
        }
    } else {
        c->cs_const = -1.0;
    }
    if (c->d.sgs && sp->sgs_model == 5) {
        // sgs_stag_util.f90:183-231 with the coefficient field
        if (sp->lasd_cs_init) { if (fill(c, F[LG_CS_OPT2], c->plane, 0, nz + 1, 0.03)) return 1; }
        else if (sp->lasd_update && lagrange_sdep(c, sp, F)) return 1;
        LasdSSArgs na;
        for (int i = 0; i < 6; ++i) { na.S[i] = c->work[i]; na.SS[i] = nullptr; }
        ProfScope ps_(c, "sgs");
        LG_LAUNCH(k_nut_field, dim3(grid1d(long(c->nx) * c->ny * nz)), dim3(kBlock), 0, c->stream, na, F[LG_CS_OPT2], c->lsq,
                  c->work[6], c->lay(), c->nx, c->ny, 1, nz + 1);
        c->launches++;
    }
    TauArgs ta;
    for (int i = 0; i < 6; ++i) ta.S[i] = c->work[i];
    ta.Nu_t = c->work[6];
    ta.T[0] = F[LG_TXX]; ta.T[1] = F[LG_TXY]; ta.T[2] = F[LG_TXZ]; ta.T[3] = F[LG_TYY]; ta.T[4] = F[LG_TYZ]; ta.T[5] = F[LG_TZZ];
    {
        ProfScope ps_(c, "sgs");
        // plane 1 on the bottom rank keeps the wall-model txz, tyz: k_tau does not write them there
        LG_LAUNCH(k_tau, dim3(grid1d(long(c->nx) * c->ny * (nz - 1))), dim3(kBlock), 0, c->stream, ta, p, c->lay(), c->nx, c->ny, 1, nz);
        c->launches++;
    }
    // txz, tyz: plane 1 of coord+1 -> plane nz of coord (sgs_stag_util.f90:437-444); tzz(nz-1) -> tzz(0) above (main.f90:194)
    if (plane_exchange(c, F[LG_TXZ] + c->plane, coord - 1, F[LG_TXZ] + c->plane * nz, coord + 1)) return 1;
    if (plane_exchange(c, F[LG_TYZ] + c->plane, coord - 1, F[LG_TYZ] + c->plane * nz, coord + 1)) return 1;
    if (plane_exchange(c, F[LG_TZZ] + c->plane * (nz - 1), coord + 1, F[LG_TZZ], coord - 1)) return 1;
    double** d = c->work + 7;
    // divstress_uv.f90:21-86
    if (spectral_deriv(c, F[LG_TXX], nullptr, d[0], nullptr)) return 1;          // dtxdx
    ddz_w(c, F[LG_TXZ], d[1]);                                                    // dtzdz
    if (spectral_deriv(c, F[LG_TYY], nullptr, nullptr, d[2])) return 1;          // dtydy2
    ddz_w(c, F[LG_TYZ], d[3]);                                                    // dtzdz2
    if (spectral_deriv(c, F[LG_TXY], nullptr, d[4], d[5])) return 1;             // dtxdx2, dtydy
    sum3(c, F[LG_DIVTX], d[0], d[5], d[1], 1, nz, 1);
    sum3(c, F[LG_DIVTY], d[4], d[2], d[3], 1, nz, 1);
    // divstress_w.f90:21-116
    if (spectral_deriv(c, F[LG_TXZ], nullptr, d[0], nullptr)) return 1;
    if (spectral_deriv(c, F[LG_TYZ], nullptr, nullptr, d[1])) return 1;
    ddz_uv(c, F[LG_TZZ], d[2]);
    sum3(c, F[LG_DIVTZ], d[0], d[1], c->bottom ? nullptr : d[2], 1, 2, 1);
    sum3(c, F[LG_DIVTZ], d[0], d[1], d[2], 2, nz, 1);
    sum3(c, F[LG_DIVTZ], d[0], d[1], c->top ? nullptr : d[2], nz, nz + 1, 0);
    return 0;
}

// ---- actuator disks, turbines.f90:465-638 on the resident fields -----------------------
int turbines_forcing(lesgo_gpu_ctx* c, double eps) {
    if (!c->turb_on) return c->fail("lesgo_gpu_turbines_init has not been called");
    const int nz = c->nz, nloc = c->turb.nloc;
    double* W = field(c, LG_W);
    double *fxa = field(c, LG_FXA), *fya = field(c, LG_FYA), *fza = field(c, LG_FZA);
    if (!W || !fxa || !fya || !fza) return 1;
    if (c->comm && c->comm->sync_planes(W, c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());   // :499
    if (nloc > 0) {
        ProfScope ps_(c, "turbines");
        const double vol = (c->d.L_x / c->nx) * (c->d.L_y / c->ny) * c->d.dz;
        LG_LAUNCH(k_turb_gather, dim3(nloc), dim3(kBlock), 0, c->stream, c->turb, field(c, LG_U), field(c, LG_V), W, c->plane, vol);
        if (c->comm && c->comm->allreduce_sum_dev(c->turb.u_d, size_t(nloc), c->stream)) return c->fail(c->comm->error());   // :553-560
        LG_LAUNCH(k_turb_update, dim3(1), dim3(64), 0, c->stream, c->turb, eps, c->turb_adm);
        LG_LAUNCH(k_turb_scatter, dim3(nloc), dim3(kBlock), 0, c->stream, c->turb, fxa, fya, c->turb_fzuv);
        c->launches += 3;
    }
    if (c->comm) {                                                 // :620-622
        ProfScope ps_(c, "halo");
        if (c->comm->sync_planes(fxa, c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
        if (c->comm->sync_planes(fya, c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
        if (c->turb_fz && c->comm->sync_planes(c->turb_fzuv, c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
    }
    if (c->turb_fz) {                                              // :623 interp_to_w_grid; all nhat(3) = 0: fza stays 0
        ProfScope ps_(c, "turbines");
        LG_LAUNCH(k_interp_w, dim3(grid1d(c->plane * nz)), dim3(kBlock), 0, c->stream, c->turb_fzuv, fza, c->plane, 1, nz + 1);
        c->launches++;
        if (c->comm && c->comm->sync_planes(fza, c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
    }
    return 0;
}

template <class T>
int upload_vec(lesgo_gpu_ctx* c, const T** dst, const std::vector<T>& h, std::vector<void*>* owner = nullptr) {
    void* p = nullptr;
    CK(cudaMalloc(&p, (h.empty() ? 1 : h.size()) * sizeof(T)));
    (owner ? *owner : c->allocs).push_back(p);
    if (!h.empty()) CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dst = static_cast<const T*>(p);
    return 0;
}

int turbines_init(lesgo_gpu_ctx* c, int nloc, const lesgo_gpu_turbine* t, int adm) {
    if (nloc < 0 || (nloc > 0 && !t)) return c->fail("lesgo_gpu_turbines_init: bad arguments");
    const int nz = c->nz;
    // a re-meshed farm (dyn_theta1/2: turbines.f90:506-515 calls turbines_nodes every step) replaces the old arrays
    if (!c->turb_allocs.empty()) {
        CK(cudaStreamSynchronize(c->stream));
        for (void* q : c->turb_allocs) cudaFree(q);
        c->turb_allocs.clear();
        c->turb_on = false;
    }
    std::vector<void*>* TA = &c->turb_allocs;
    std::vector<int> start(nloc + 1, 0), owner;
    std::vector<long> off;
    std::vector<double> ind, nhat(3 * size_t(nloc)), Ct(nloc), dia(nloc), M(nloc), udT(nloc);
    bool fz = false;
    for (int s = 0; s < nloc; ++s) {
        if (t[s].num_nodes < 0 || (t[s].num_nodes > 0 && (!t[s].nodes || !t[s].ind)))
            return c->fail("lesgo_gpu_turbines_init: turbine without node list");
        for (int l = 0; l < t[s].num_nodes; ++l) {
            const int i = t[s].nodes[3 * l], j = t[s].nodes[3 * l + 1], k = t[s].nodes[3 * l + 2];
            if (i < 1 || i > c->nx || j < 1 || j > c->ny || k < 1 || k > nz - 1)
                return c->fail("lesgo_gpu_turbines_init: node outside 1:nx, 1:ny, 1:nz-1 (turbines.f90:423-425)");
            off.push_back(c->lay().at(k, j - 1, i - 1));
            ind.push_back(t[s].ind[l]);
        }
        start[s + 1] = int(off.size());
        for (int q = 0; q < 3; ++q) nhat[3 * s + q] = t[s].nhat[q];
        if (t[s].nhat[2] != 0.0) fz = true;
        Ct[s] = t[s].Ct_prime; dia[s] = t[s].dia; M[s] = t[s].M; udT[s] = t[s].u_d_T;
    }
    // the reference ASSIGNS the force node by node in disk order (turbines.f90:599-606): where disks overlap the
    // last one wins, so only the last entry of a grid point scatters
    owner.assign(off.size(), 1);
    {
        std::vector<size_t> order(off.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return off[a] < off[b]; });
        for (size_t i = 0; i + 1 < order.size(); ++i)
            if (off[order[i]] == off[order[i + 1]]) owner[order[i]] = 0;
    }
    TurbSet ts;
    ts.nloc = nloc;
    const double *d_udT = nullptr, *d_zero = nullptr;
    if (upload_vec(c, &ts.start, start, TA) || upload_vec(c, &ts.off, off, TA) || upload_vec(c, &ts.ind, ind, TA) ||
        upload_vec(c, &ts.owner, owner, TA) || upload_vec(c, &ts.nhat, nhat, TA) || upload_vec(c, &ts.Ct_prime, Ct, TA) ||
        upload_vec(c, &ts.dia, dia, TA) || upload_vec(c, &ts.M, M, TA) || upload_vec(c, &d_udT, udT, TA)) return 1;
    std::vector<double> z(nloc, 0.0);
    if (upload_vec(c, &d_zero, z, TA)) return 1;
    ts.u_d_T = const_cast<double*>(d_udT);
    ts.u_d = const_cast<double*>(d_zero);
    if (upload_vec(c, &d_zero, z, TA)) return 1;
    ts.f_n = const_cast<double*>(d_zero);
    ts.ind_t = nullptr; ts.e_theta = nullptr; ts.tip_speed_ratio = 1.0;
    c->turb_nodes.assign(start.begin(), start.end());
    c->turb = ts; c->turb_adm = adm; c->turb_fz = fz;
    // forcing.f90:103-105: the force fields are zero away from the disks
    const size_t nfield = size_t(c->plane) * (nz + 1);
    if (dev_alloc(c, &c->turb_fzuv, nfield)) return 1;
    double* f[4] = {field(c, LG_FXA), field(c, LG_FYA), field(c, LG_FZA), c->turb_fzuv};
    for (int i = 0; i < 4; ++i) {
        if (!f[i]) return 1;
        CK(cudaMemsetAsync(f[i], 0, nfield * sizeof(double), c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    c->turb_on = true;
    return 0;
}

// ---- one timestep on the resident fields: main.f90:155-344 ----------------------------------------
int step(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp) {
    const int nz = c->nz;
    double* F[LG_NFIELDS];
    const bool lasd = sp->mode == 1 && c->d.sgs && sp->sgs_model == 5;
    for (int i = 0; i < LG_NFIELDS; ++i) {
        const bool want = i < LG_F_LM || (i < LG_FXA ? lasd : c->turb_on);
        F[i] = want ? field(c, i) : nullptr;
        if (!F[i] && want) return 1;
    }
    if (sp->turbines && !c->turb_on) return c->fail("lesgo_gpu_step: turbines = 1 needs lesgo_gpu_turbines_init");
    if (sp->turbines && c->chunk != 0) return c->fail("lesgo_gpu_step: turbines need the un-chunked step (LESGO_CHUNK=0)");
    if (sp->mode != 0 && sp->mode != 1) return c->fail("lesgo_gpu_step: mode must be 0 (core) or 1 (full)");
    // :155-157  RHS*_f = RHS*: the two sets trade places instead of being copied (convec
    // rewrites every valid plane of RHS* below, main.f90:207-214)
    std::swap(c->fields[LG_RHSX], c->fields[LG_RHSX_F]); std::swap(F[LG_RHSX], F[LG_RHSX_F]);
    std::swap(c->fields[LG_RHSY], c->fields[LG_RHSY_F]); std::swap(F[LG_RHSY], F[LG_RHSY_F]);
    std::swap(c->fields[LG_RHSZ], c->fields[LG_RHSZ_F]); std::swap(F[LG_RHSZ], F[LG_RHSZ_F]);
    // :161-172
    // spectral reuse: filt_da also emits the 3/2-grid y transforms of u, v, w for convec (LESGO_REUSE=0: off)
    const bool reuse = reuse_enabled() && c->chunk == 0 && !HP(c);
    if (reuse)
        for (int i = 0; i < 3; ++i)
            if (dev_alloc(c, &c->bb[i], size_t(c->plane_bi) * (nz + 1))) return 1;
    if (batch_enabled() && c->chunk == 0 && !HP(c)) {
        if (spectral_deriv_uvw(c, F, reuse)) return 1;
    } else {
        if (spectral_deriv(c, F[LG_U], F[LG_U], F[LG_DUDX], F[LG_DUDY], reuse ? c->bb[0] : nullptr)) return 1;
        if (spectral_deriv(c, F[LG_V], F[LG_V], F[LG_DVDX], F[LG_DVDY], reuse ? c->bb[1] : nullptr)) return 1;
        if (spectral_deriv(c, F[LG_W], F[LG_W], F[LG_DWDX], F[LG_DWDY], reuse ? c->bb[2] : nullptr)) return 1;
    }
    if (ddz_uv(c, F[LG_U], F[LG_DUDZ])) return 1;
    if (ddz_uv(c, F[LG_V], F[LG_DVDZ])) return 1;
    if (ddz_w(c, F[LG_W], F[LG_DWDZ])) return 1;
    // wallstress (:182-184); sgs_stag, tzz halo, divstress_uv/w (:189-203) in the full mode
    if (sp->mode == 1 && build_sgs_tables(c, sp)) return 1;
    if (wallstress(c, sp, F, sp->mode == 1)) return 1;
    if (sp->mode == 1 && sgs_and_divstress(c, sp, F)) return 1;
    // :254 forcing_applied: the disks see the velocities of time level n, which the fused epilogue below
    // advances -- so the forcing runs first; :264-266 RHS += f rides in that epilogue
    if (sp->turbines && turbines_forcing(c, sp->turbines_eps)) return 1;
    // :207 convec, with :211-214/229-232 (RHS assembly), :273-280 (Euler start) and :287-296 (AB2)
    // fused into the epilogue of its last x pass
    {
        Fuse fz;
        if (sp->turbines) { fz.fa[0] = F[LG_FXA]; fz.fa[1] = F[LG_FYA]; fz.fa[2] = F[LG_FZA]; fz.kfa = nz - 1; }
        fz.mode = 1; fz.first_step = sp->first_step ? 1 : 0; fz.dt = sp->dt; fz.t1 = sp->tadv1; fz.t2 = sp->tadv2;
        fz.divt[0] = F[LG_DIVTX]; fz.divt[1] = F[LG_DIVTY]; fz.divt[2] = F[LG_DIVTZ];
        fz.rhs_f[0] = F[LG_RHSX_F]; fz.rhs_f[1] = F[LG_RHSY_F]; fz.rhs_f[2] = F[LG_RHSZ_F];
        fz.u[0] = F[LG_U]; fz.u[1] = F[LG_V]; fz.u[2] = F[LG_W];
        fz.force[0] = sp->mean_p_force_x; fz.force[1] = sp->mean_p_force_y; fz.force[2] = 0.0;
        fz.kmax[0] = nz - 1; fz.kmax[1] = nz - 1; fz.kmax[2] = c->top ? nz : nz - 1;
        // convec reads u, v, w only in its first passes (to the 3/2 grid), all finished before the
        // epilogue of the last pass updates them -- but only when the whole slab is one chunk
        const Fuse* use = c->chunk == 0 ? &fz : nullptr;
        if (convec(c, F[LG_U], F[LG_V], F[LG_W], F[LG_DUDY], F[LG_DUDZ], F[LG_DVDX], F[LG_DVDZ], F[LG_DWDX],
                   F[LG_DWDY], F[LG_RHSX], F[LG_RHSY], F[LG_RHSZ], use, reuse)) return 1;
        if (!use) {
            const int kw = c->top ? nz + 1 : nz;
            if (glue_fused(c, F_RHS_AB2, F[LG_RHSX], F[LG_DIVTX], F[LG_RHSX_F], F[LG_U], 1, nz, 0, fz.first_step, sp->mean_p_force_x, sp->dt, sp->tadv1, sp->tadv2)) return 1;
            if (glue_fused(c, F_RHS_AB2, F[LG_RHSY], F[LG_DIVTY], F[LG_RHSY_F], F[LG_V], 1, nz, 0, fz.first_step, sp->mean_p_force_y, sp->dt, sp->tadv1, sp->tadv2)) return 1;
            if (glue_fused(c, F_RHS_AB2, F[LG_RHSZ], F[LG_DIVTZ], F[LG_RHSZ_F], F[LG_W], 1, kw, 0, fz.first_step, 0.0, sp->dt, sp->tadv1, sp->tadv2)) return 1;
        }
    }
    // :299-308
    {
        FillGroup fg_(c);
        fill(c, F[LG_U], c->plane, 0, 1, kBogus); fill(c, F[LG_V], c->plane, 0, 1, kBogus); fill(c, F[LG_W], c->plane, 0, 1, kBogus);
        fill(c, F[LG_U], c->plane, nz, nz + 1, kBogus); fill(c, F[LG_V], c->plane, nz, nz + 1, kBogus);
        if (!c->top) fill(c, F[LG_W], c->plane, nz, nz + 1, kBogus);
    }
    // :317 press_stag_array, with :321-326 (RHS -= grad p) and project (forcing.f90:171-207) fused into
    // the epilogues of its last passes.  press reads u, v, w only in its first pass.
    {
        Fuse fz;
        fz.mode = 2; fz.dt = sp->dt; fz.t1 = sp->tadv1;
        fz.rhs_f[0] = F[LG_RHSX]; fz.rhs_f[1] = F[LG_RHSY];      // mode 2 updates RHS in place
        fz.u[0] = F[LG_U]; fz.u[1] = F[LG_V];
        fz.rhsz = F[LG_RHSZ]; fz.w = F[LG_W];
        fz.kproj = c->bottom ? 2 : 1; fz.kproj_end = nz;
        if (press(c, F[LG_U], F[LG_V], F[LG_W], F[LG_DIVTZ], sp->dt, sp->tadv1, F[LG_P], F[LG_DPDX], F[LG_DPDY],
                  F[LG_DPDZ], &fz)) return 1;
    }
    if (c->comm) {
        if (c->comm->sync_planes(F[LG_U], c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
        if (c->comm->sync_planes(F[LG_V], c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
        if (c->comm->sync_planes(F[LG_W], c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
    }
    if (c->top) {
        if (c->d.ubc_mom == 0) {
            CK(cudaMemcpyAsync(F[LG_U] + c->plane * nz, F[LG_U] + c->plane * (nz - 1), c->plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            CK(cudaMemcpyAsync(F[LG_V] + c->plane * nz, F[LG_V] + c->plane * (nz - 1), c->plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
        fill(c, F[LG_W], c->plane, nz, nz + 1, 0.0);
    }
    if (c->bottom) fill(c, F[LG_W], c->plane, 1, 2, 0.0);
    return 0;
}

}  // namespace

// =====================================================================================================
// every entry point makes the context's device current for the calling thread
#define ENTER(c) do { if (c) cudaSetDevice((c)->device); } while (0)

extern "C" {

const char* lesgo_gpu_last_error(const lesgo_gpu_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }

int lesgo_gpu_create(const lesgo_gpu_dims* d, lesgo_gpu_ctx** out) {
    if (!d || !out) { g_err = "null argument"; return 1; }
    *out = nullptr;
    lesgo_gpu_ctx* c = new lesgo_gpu_ctx;
    c->d = *d;
    auto bail = [&](const std::string& m) { g_err = m; delete c; return 1; };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return bail("no CUDA device: liblesgo_cuda has no CPU fallback");
    // device >= 0: that ordinal; -1: the calling thread's current device; <= -2: node-local rank r = -2 - device,
    // mapped to ordinal r mod (device count) -- what the Fortran shim passes under mpirun, where every rank is its
    // own process whose current device would otherwise be 0
    int want = d->device;
    if (want <= -2) want = (-2 - want) % ndev;
    if (want >= 0) {
        if (want >= ndev) return bail("device ordinal out of range");
        if (cudaSetDevice(want) != cudaSuccess) return bail("cudaSetDevice failed");
    }
    cudaGetDevice(&c->device);
    c->d.device = c->device;
    if (d->nx < 16 || d->ny < 16 || d->nz < 2 || (d->nx % 4) || (d->ny % 4)) return bail("bad grid size");
    if (!size_supported(d->nx) || !size_supported(d->ny)) return bail("nx/ny not in the supported FFT size list (sizes.h)");
    if (d->nproc < 1 || d->coord < 0 || d->coord >= d->nproc) return bail("bad nproc/coord");
    if (d->nz_tot != (d->nz - 1) * d->nproc + 1) return bail("nz_tot must equal (nz-1)*nproc+1 (input_util.f90:200)");
    c->nx = d->nx; c->ny = d->ny; c->nz = d->nz; c->nzt = d->nz_tot;
    c->lh = c->nx / 2 + 1; c->ld = 2 * c->lh;
    c->nx2 = 3 * c->nx / 2; c->ny2 = 3 * c->ny / 2;
    c->lh_big = c->nx2 / 2 + 1; c->ld_big = 2 * c->lh_big;
    c->plane = long(c->ld) * c->ny; c->plane_big = long(c->ld_big) * c->ny2; c->plane_bi = long(c->ld) * c->ny2;
    c->bottom = d->coord == 0; c->top = d->coord == d->nproc - 1;
    c->jzLo = d->sgs ? 2 : 1;
    const double pi = 3.14159265358979323846;   // param.f90 pi
    c->kxs = 2.0 * pi / d->L_x; c->kys = 2.0 * pi / d->L_y;
    {
        // planes per pipelined chunk (LESGO_CHUNK_PLANES; default 0 = every pass covers the whole
        // slab).  Measured on B200 at 512x512x257 (profiles/r1_chunk_sweep.md): chunks of 4/8/16/32
        // planes cost 59.6/44.6/39.5/35.1 ms per step against 33.5 ms un-chunked -- the passes are
        // latency-, not DRAM-bound, so L2 residency does not pay for the extra launches.
        int ch = 0;
        if (const char* e = std::getenv("LESGO_CHUNK_PLANES")) ch = std::atoi(e);
        c->chunk = (ch <= 0 || ch >= c->nz + 1) ? 0 : ch;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("stream create failed");
    c->own_stream = true;
    int rc = 0;
    rc |= make_stage_twiddles(c, &c->Wx, c->nx / 2);
    rc |= make_half_twiddles(c, &c->Whx, c->nx / 2);
    rc |= make_stage_twiddles(c, &c->Wxb, c->nx2 / 2);
    rc |= make_half_twiddles(c, &c->Whxb, c->nx2 / 2);
    rc |= make_stage_twiddles(c, &c->Wy, c->ny, true);
    rc |= make_stage_twiddles(c, &c->Wyb, c->ny2, true);
    if (rc) { std::string m = c->err; return bail(m); }
    *out = c;
    lesgo_gpu_fftw_bind(c, &c->d);      // the dfftw_* symbols serve this context's plans (fftw_shim.cu)
    return 0;
}

int lesgo_gpu_destroy(lesgo_gpu_ctx* c) {
    ENTER(c);
    if (!c) return 0;
    if (lesgo_gpu_fftw_bound() == c) lesgo_gpu_fftw_bind(nullptr, nullptr);
    cudaStreamSynchronize(c->stream);
    if (c->comm) { delete c->comm; c->comm = nullptr; }
#ifndef LESGO_EMUL
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
#endif
    for (void* p : c->allocs) cudaFree(p);
    for (void* p : c->turb_allocs) cudaFree(p);
    for (double* p : c->staging) if (p) cudaFree(p);
    if (c->red_host) cudaFreeHost(c->red_host);
    if (c->s_in) { cudaStreamDestroy(c->s_in); cudaStreamDestroy(c->s_out); }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

// Page-lock a host array the caller keeps passing to the per-routine entry points (the Fortran module arrays of
// sim_param are ordinary pageable allocations): staged copies then run at full PCIe rate and truly overlap with
// the kernels of the chunk pipeline.  Idempotent; unregister before the array is deallocated.
int lesgo_gpu_host_register(lesgo_gpu_ctx* c, void* host, size_t bytes) {
    ENTER(c);
    if (!c || !host || bytes == 0) return 1;
#ifndef LESGO_EMUL
    cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return c->fail(std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
#endif
    return 0;
}

int lesgo_gpu_host_unregister(lesgo_gpu_ctx* c, void* host) {
    ENTER(c);
    if (!c || !host) return 1;
#ifndef LESGO_EMUL
    cudaError_t e = cudaHostUnregister(host);
    if (e != cudaSuccess && e != cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return c->fail(std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
    cudaGetLastError();
#endif
    return 0;
}

int lesgo_gpu_set_stream(lesgo_gpu_ctx* c, void* s) {
    ENTER(c);
    if (!c) return 1;
    cudaStreamSynchronize(c->stream);
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = static_cast<cudaStream_t>(s);
    return 0;
}

int lesgo_gpu_synchronize(lesgo_gpu_ctx* c) {
    ENTER(c);
    if (!c) return 1;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return 0;
}

long lesgo_gpu_launch_count(const lesgo_gpu_ctx* c) { return c ? c->launches : 0; }

int lesgo_gpu_profile(lesgo_gpu_ctx* c, int enable, char* report, int report_len) {
    ENTER(c);
    // enable = 1/0 switches per-launch event timing; a non-NULL report receives
    // "label count total_ms\n" lines for everything recorded so far and clears the records.
    if (!c) return 1;
#ifndef LESGO_EMUL
    if (report && report_len > 0) {
        CK(cudaStreamSynchronize(c->stream));
        struct Acc { const char* l; int n; double ms; };
        std::vector<Acc> acc;
        for (auto& r : c->prof_recs) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.a, r.b);
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
            bool found = false;
            for (auto& a : acc) if (std::strcmp(a.l, r.label) == 0) { a.n++; a.ms += ms; found = true; break; }
            if (!found) acc.push_back(Acc{r.label, 1, double(ms)});
        }
        c->prof_recs.clear();
        std::string out;
        char line[160];
        for (auto& a : acc) { std::snprintf(line, sizeof line, "%s %d %.6f\n", a.l, a.n, a.ms); out += line; }
        std::snprintf(report, size_t(report_len), "%s", out.c_str());
    }
#else
    if (report && report_len > 0) report[0] = 0;
#endif
    c->prof = enable != 0;
    return 0;
}

int lesgo_gpu_wavenumbers(lesgo_gpu_ctx* c, double* kx, double* ky, double* k2) {
    // fft.f90:130-160
    if (!c) return 1;
    for (int jy = 0; jy < c->ny; ++jy)
        for (int jx = 0; jx < c->lh; ++jx) {
            double x = double(jx), y = double(((jy + c->ny / 2) % c->ny) - c->ny / 2);
            if (jx == c->lh - 1 || jy == c->ny / 2) { x = 0.0; y = 0.0; }
            x = c->kxs * x; y = c->kys * y;
            const long o = long(jy) * c->lh + jx;
            if (kx) kx[o] = x;
            if (ky) ky[o] = y;
            if (k2) k2[o] = x * x + y * y;
        }
    return 0;
}

#define NFIELD (size_t(c->plane) * (c->nz + 1))

int lesgo_gpu_filt_da(lesgo_gpu_ctx* c, double* f, double* dfdx, double* dfdy) {
    ENTER(c);
    if (!c || !f || !dfdx || !dfdy) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, true);
    double* dx = st.in(dfdx, NFIELD, false, true);
    double* dy = st.in(dfdy, NFIELD, false, true);
    if (!df || !dx || !dy) return 1;
    st.begin();
    if (spectral_deriv(c, df, df, dx, dy)) return 1;
    return st.finish();
}

int lesgo_gpu_ddx(lesgo_gpu_ctx* c, const double* f, double* dfdx) {
    ENTER(c);
    if (!c || !f || !dfdx) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, false);
    double* dx = st.in(dfdx, NFIELD, false, true);
    if (!df || !dx) return 1;
    st.begin();
    if (spectral_deriv(c, df, nullptr, dx, nullptr)) return 1;
    return st.finish();
}

int lesgo_gpu_ddy(lesgo_gpu_ctx* c, const double* f, double* dfdy) {
    ENTER(c);
    if (!c || !f || !dfdy) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, false);
    double* dy = st.in(dfdy, NFIELD, false, true);
    if (!df || !dy) return 1;
    st.begin();
    if (spectral_deriv(c, df, nullptr, nullptr, dy)) return 1;
    return st.finish();
}

int lesgo_gpu_ddxy(lesgo_gpu_ctx* c, const double* f, double* dfdx, double* dfdy) {
    ENTER(c);
    if (!c || !f || !dfdx || !dfdy) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, false);
    double* dx = st.in(dfdx, NFIELD, false, true);
    double* dy = st.in(dfdy, NFIELD, false, true);
    if (!df || !dx || !dy) return 1;
    st.begin();
    if (spectral_deriv(c, df, nullptr, dx, dy)) return 1;
    return st.finish();
}

int lesgo_gpu_ddz_uv(lesgo_gpu_ctx* c, const double* f, double* dfdz) {
    ENTER(c);
    if (!c || !f || !dfdz) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, false);
    double* dz = st.in(dfdz, NFIELD, false, true);
    if (!df || !dz) return 1;
    st.begin();
    if (ddz_uv(c, df, dz)) return 1;
    return st.finish();
}

int lesgo_gpu_ddz_w(lesgo_gpu_ctx* c, const double* f, double* dfdz) {
    ENTER(c);
    if (!c || !f || !dfdz) return 1;
    Staged st(c, true);
    double* df = st.in(f, NFIELD, true, false);
    double* dz = st.in(dfdz, NFIELD, false, true);
    if (!df || !dz) return 1;
    st.begin();
    if (ddz_w(c, df, dz)) return 1;
    return st.finish();
}

int lesgo_gpu_convec(lesgo_gpu_ctx* c, const double* u, const double* v, const double* w, const double* dudy,
                     const double* dudz, const double* dvdx, const double* dvdz, const double* dwdx,
                     const double* dwdy, double* RHSx, double* RHSy, double* RHSz) {
    ENTER(c);
    if (!c) return 1;
    Staged st(c, true);
    const double* in[9] = {u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy};
    double* din[9];
    for (int i = 0; i < 9; ++i) {
        if (!in[i]) return c->fail("convec: null input");
        din[i] = st.in(in[i], NFIELD, true, false);
        if (!din[i]) return 1;
    }
    double* ox = st.in(RHSx, NFIELD, false, true);
    double* oy = st.in(RHSy, NFIELD, false, true);
    double* oz = st.in(RHSz, NFIELD, false, true);
    if (!ox || !oy || !oz) return c->fail("convec: null output");
    st.begin();
    if (convec(c, din[0], din[1], din[2], din[3], din[4], din[5], din[6], din[7], din[8], ox, oy, oz)) return 1;
    return st.finish();
}

int lesgo_gpu_press_stag_array(lesgo_gpu_ctx* c, const double* u, const double* v, const double* w,
                               const double* divtz, double dt, double tadv1, double* p, double* dpdx,
                               double* dpdy, double* dpdz) {
    ENTER(c);
    if (!c || !u || !v || !w || !divtz || !p || !dpdx || !dpdy || !dpdz) return 1;
    Staged st(c, true);
    double* du = st.in(u, NFIELD, true, false);
    double* dv = st.in(v, NFIELD, true, false);
    double* dw = st.in(w, NFIELD, true, false);
    double* dd = st.in(divtz, NFIELD, true, false);
    double* op = st.in(p, NFIELD, false, true);
    // dpdx, dpdy, dpdz are (1:nz) in the reference: plane 0 of the passed address is not theirs
    double* ox = st.in(dpdx, NFIELD, false, true, size_t(c->plane));
    double* oy = st.in(dpdy, NFIELD, false, true, size_t(c->plane));
    double* oz = st.in(dpdz, NFIELD, false, true, size_t(c->plane));
    if (!du || !dv || !dw || !dd || !op || !ox || !oy || !oz) return 1;
    st.begin();
    if (press(c, du, dv, dw, dd, dt, tadv1, op, ox, oy, oz)) return 1;
    return st.finish();
}

int lesgo_gpu_fft_r2c(lesgo_gpu_ctx* c, const double* in, double* out, int nplanes, int bigg) {
    ENTER(c);
    if (!c || !in || !out || nplanes < 1) return 1;
    const bool b = bigg != 0;
    const long pl = b ? c->plane_big : c->plane;
    const int row = b ? c->ld_big : c->ld, nyr = b ? c->ny2 : c->ny, nxr = b ? c->nx2 : c->nx;
    Staged st(c);
    double* di = st.in(in, size_t(pl) * nplanes, true, false);
    double* dout = st.in(out, size_t(pl) * nplanes, false, true);
    if (!di || !dout) return 1;
    ProScale ps; ps.src[0] = di; ps.lay = Lay{pl, row}; ps.scale = 1.0;
    double* d[1] = {dout};
    if (xfwd(c, b, ps, 1, d, pl, row, nxr / 2, nyr, 0, nplanes, 2)) return 1;
    YArgs a = yargs(c, pl, row, pl, row, nxr / 2 + 1, 0);
    a.keep_nyq_row = 1;
    a.fld[0].src = dout; a.fld[0].out[0] = YOutSpec{dout, Y_COPY};
    if (ypass(c, nyr, 0, a, 1, 0, nplanes)) return 1;
    return st.finish();
}

int lesgo_gpu_fft_c2r(lesgo_gpu_ctx* c, const double* in, double* out, int nplanes, int bigg) {
    ENTER(c);
    if (!c || !in || !out || nplanes < 1) return 1;
    const bool b = bigg != 0;
    const long pl = b ? c->plane_big : c->plane;
    const int row = b ? c->ld_big : c->ld, nyr = b ? c->ny2 : c->ny, nxr = b ? c->nx2 : c->nx;
    Staged st(c);
    double* di = st.in(in, size_t(pl) * nplanes, true, false);
    double* dout = st.in(out, size_t(pl) * nplanes, false, true);
    if (!di || !dout) return 1;
    // y inverse into a scratch spectrum, then x inverse
    double* tmp = nullptr;
    if (b) { if (need_big(c, 1)) return 1; }
    else if (need_small(c, 1)) return 1;
    if (nplanes > c->nz + 1) return c->fail("fft_c2r: at most nz+1 planes per call");
    tmp = b ? c->big[0] : c->sa[0];
    YArgs a = yargs(c, pl, row, pl, row, nxr / 2 + 1, 0);
    a.keep_nyq_row = 1;
    a.fld[0].src = di; a.fld[0].out[0] = YOutSpec{tmp, Y_COPY};
    if (ypass(c, 0, nyr, a, 1, 0, nplanes)) return 1;
    const double* s0[1] = {tmp};
    double* o0[1] = {dout};
    if (xinv(c, b, s0, pl, row, nxr / 2 + 1, 1, o0, Lay{pl, row}, nyr, 0, nplanes, 0)) return 1;
    return st.finish();
}

double* lesgo_gpu_field_ptr(lesgo_gpu_ctx* c, int id) {
    ENTER(c); return c ? field(c, id) : nullptr; }

int lesgo_gpu_upload(lesgo_gpu_ctx* c, int id, const double* host) {
    ENTER(c);
    if (!c || !host) return 1;
    double* f = field(c, id);
    if (!f) return c->fail("bad field id");
    if (id == LG_CS_OPT2) c->cs_const = -1.0;
    CK(cudaMemcpyAsync(f, host, NFIELD * sizeof(double), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int lesgo_gpu_download(lesgo_gpu_ctx* c, int id, double* host) {
    ENTER(c);
    if (!c || !host) return 1;
    double* f = field(c, id);
    if (!f) return c->fail("bad field id");
    CK(cudaMemcpyAsync(host, f, NFIELD * sizeof(double), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int lesgo_gpu_step(lesgo_gpu_ctx* c, const lesgo_gpu_step_params* sp) {
    ENTER(c);
    if (!c || !sp) return 1;
    if (step(c, sp)) return 1;
    CK(cudaGetLastError());
    return 0;
}

// ---- restart file of the resident state -------------------------------------------------
// One Fortran sequential unformatted record (io.f90:1204-1211 / initial.f90:226-239) holding planes 1:nz of
// u, v, w, RHSx, RHSy, RHSz, Cs_opt2, F_LM, F_MM, F_QN, F_NN, native byte order.  Records longer than
// 2 GiB - 9 bytes use gfortran's subrecord convention: 4-byte signed byte counts before and after every
// subrecord, the leading one negative when another subrecord follows, the trailing one negative when one
// precedes (LESGO_SUBRECORD_MAX overrides the limit so tests can exercise it with small files).
namespace {
const int kCkptFields[11] = {LG_U, LG_V, LG_W, LG_RHSX, LG_RHSY, LG_RHSZ, LG_CS_OPT2, LG_F_LM, LG_F_MM, LG_F_QN, LG_F_NN};
long long subrecord_max() {
    const char* e = std::getenv("LESGO_SUBRECORD_MAX");
    const long long v = e ? std::atoll(e) : 0;
    return v > 0 ? v : 2147483639LL;
}
// streams `total` bytes of the record between the file and `io(ptr, n)`-sized pieces of payload
struct RecordIO {
    FILE* f; bool writing; long long total, done = 0, sub_left = 0, maxsub; bool first = true;
    std::string err;
    RecordIO(FILE* f_, bool w, long long tot) : f(f_), writing(w), total(tot), maxsub(subrecord_max()) {}
    bool marker(int v) {
        if (writing) return std::fwrite(&v, 4, 1, f) == 1;
        int r = 0;
        if (std::fread(&r, 4, 1, f) != 1) { err = "unexpected end of file"; return false; }
        if (r != v) { err = "record length " + std::to_string(r) + " where " + std::to_string(v) + " was expected (other grid or byte order?)"; return false; }
        return true;
    }
    bool open_sub() {
        const long long left = total - done;
        sub_left = left < maxsub ? left : maxsub;
        cur = int(sub_left);
        return marker(left > sub_left ? -cur : cur);
    }
    bool close_sub() { const bool ok = marker(first ? cur : -cur); first = false; return ok; }
    int cur = 0;
    bool xfer(char* p, long long n) {
        while (n > 0) {
            if (sub_left == 0 && !open_sub()) return false;
            const long long m = n < sub_left ? n : sub_left;
            const size_t got = writing ? std::fwrite(p, 1, size_t(m), f) : std::fread(p, 1, size_t(m), f);
            if (got != size_t(m)) { err = writing ? "write failed" : "unexpected end of file"; return false; }
            p += m; n -= m; done += m; sub_left -= m;
            if (sub_left == 0 && !close_sub()) return false;
        }
        return true;
    }
};
int checkpoint_io(lesgo_gpu_ctx* c, const char* fname, bool writing) {
    if (!fname) return c->fail("checkpoint: no file name");
    FILE* f = std::fopen(fname, writing ? "wb" : "rb");
    if (!f) return c->fail(std::string("checkpoint: cannot open ") + fname);
    const long long per_field = (long long)c->plane * c->nz * sizeof(double);
    RecordIO rec(f, writing, 11 * per_field);
    const size_t chunk_planes = 8;
    const size_t chunk = size_t(c->plane) * chunk_planes;
    double* host = nullptr;
    if (cudaMallocHost(reinterpret_cast<void**>(&host), chunk * sizeof(double)) != cudaSuccess) { std::fclose(f); return c->fail("checkpoint: pinned buffer"); }
    int rc = 0;
    for (int q = 0; q < 11 && !rc; ++q) {
        double* d = field(c, kCkptFields[q]);
        if (!d) { rc = 1; break; }
        for (size_t k = 1; k <= size_t(c->nz) && !rc; k += chunk_planes) {
            const size_t np = (k + chunk_planes <= size_t(c->nz) + 1) ? chunk_planes : size_t(c->nz) + 1 - k;
            const size_t nb = np * size_t(c->plane) * sizeof(double);
            double* dp = d + size_t(c->plane) * k;
            if (writing) {
                if (cudaMemcpyAsync(host, dp, nb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = c->fail("checkpoint: device read failed"); break; }
                if (!rec.xfer(reinterpret_cast<char*>(host), (long long)nb)) rc = c->fail("checkpoint " + std::string(fname) + ": " + rec.err);
            } else {
                if (!rec.xfer(reinterpret_cast<char*>(host), (long long)nb)) { rc = c->fail("checkpoint " + std::string(fname) + ": " + rec.err); break; }
                if (cudaMemcpyAsync(dp, host, nb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) rc = c->fail("checkpoint: device write failed");
            }
        }
    }
    if (!rc && !writing) {
        char extra;
        if (std::fread(&extra, 1, 1, f) == 1) rc = c->fail("checkpoint " + std::string(fname) + ": trailing data after the record (other grid?)");
    }
    cudaFreeHost(host);
    if (std::fclose(f) != 0 && !rc) rc = c->fail("checkpoint: close failed");
    return rc;
}
}  // namespace

int lesgo_gpu_checkpoint_write(lesgo_gpu_ctx* c, const char* fname) {
    ENTER(c);
    if (!c) return 1;
    return checkpoint_io(c, fname, true);
}

int lesgo_gpu_checkpoint_read(lesgo_gpu_ctx* c, const char* fname) {
    ENTER(c);
    if (!c) return 1;
    c->cs_const = -1.0;                                            // the file's Cs_opt2 replaces the field
    return checkpoint_io(c, fname, false);
}

// ---- tavg%compute (time_average.f90:176-320) on the resident fields -------------------
int lesgo_gpu_tavg_compute(lesgo_gpu_ctx* c, double dt) {
    ENTER(c);
    if (!c) return 1;
    const int nz = c->nz;
    const size_t nfield = size_t(c->plane) * (nz + 1);
    for (int i = 0; i < TA_N; ++i)
        if (dev_alloc(c, &c->tavg_acc[i], nfield)) return 1;
    for (int i = 0; i < 5; ++i)
        if (dev_alloc(c, &c->tavg_tmp[i], nfield)) return 1;
    const bool forces = c->turb_on;
    TavgInterpArgs ia;
    ia.u = field(c, LG_U); ia.v = field(c, LG_V); ia.w = field(c, LG_W); ia.dvdx = field(c, LG_DVDX); ia.dudy = field(c, LG_DUDY);
    ia.fza = forces ? field(c, LG_FZA) : nullptr;
    ia.w_uv = c->tavg_tmp[0]; ia.u_w = c->tavg_tmp[1]; ia.v_w = c->tavg_tmp[2]; ia.vortz = c->tavg_tmp[3]; ia.fza_uv = c->tavg_tmp[4];
    ia.nz = nz; ia.top = c->top;
    {
        ProfScope ps_(c, "tavg");
        LG_LAUNCH(k_tavg_interp, dim3(grid1d(long(c->nx) * c->ny * nz)), dim3(kBlock), 0, c->stream, ia, c->lay(), c->nx, c->ny);
        c->launches++;
    }
    if (c->comm)                                                   // the syncs inside interp_to_uv/w_grid, functions.f90:83-88,131-137
        for (int i = 0; i < (forces ? 5 : 4); ++i)
            if (c->comm->sync_planes(c->tavg_tmp[i], c->plane, nz, 3, c->stream)) return c->fail(c->comm->error());
    TavgArgs a;
    a.u = ia.u; a.v = ia.v; a.w = ia.w; a.p = field(c, LG_P);
    a.txx = field(c, LG_TXX); a.tyy = field(c, LG_TYY); a.tzz = field(c, LG_TZZ); a.txy = field(c, LG_TXY);
    a.txz = field(c, LG_TXZ); a.tyz = field(c, LG_TYZ);
    a.dudz = field(c, LG_DUDZ); a.dvdz = field(c, LG_DVDZ); a.dwdx = field(c, LG_DWDX); a.dwdy = field(c, LG_DWDY);
    a.cs = field(c, LG_CS_OPT2);
    a.fxa = forces ? field(c, LG_FXA) : nullptr; a.fya = forces ? field(c, LG_FYA) : nullptr;
    a.w_uv = ia.w_uv; a.u_w = ia.u_w; a.v_w = ia.v_w; a.vortz = ia.vortz; a.fza_uv = ia.fza_uv;
    for (int i = 0; i < TA_N; ++i) a.acc[i] = c->tavg_acc[i];
    a.dt = dt; a.nz = nz; a.bottom = c->bottom; a.top = c->top; a.lbc_mom = c->d.lbc_mom; a.ubc_mom = c->d.ubc_mom;
    a.forces = forces ? 1 : 0;
    if (!a.p || !a.txx || !a.cs) return 1;
    {
        ProfScope ps_(c, "tavg");
        LG_LAUNCH(k_tavg_accumulate, dim3(grid1d(long(c->nx) * c->ny * (nz + 1))), dim3(kBlock), 0, c->stream, a, c->lay(), c->nx, c->ny);
        c->launches++;
    }
    c->tavg_time += dt;                                            // :313
    return 0;
}

int lesgo_gpu_tavg_download(lesgo_gpu_ctx* c, int which, double* host, double* total_time) {
    ENTER(c);
    if (!c) return 1;
    if (which < 0 || which >= TA_N) return c->fail("lesgo_gpu_tavg_download: bad accumulator id");
    if (!c->tavg_acc[which]) return c->fail("lesgo_gpu_tavg_download: lesgo_gpu_tavg_compute has not been called");
    if (host) {
        // device (ld, ny, 0:nz) -> host (nx, ny, lbz:nz) as tavg_t holds it (time_average.f90:98-131)
        CK(cudaMemcpy2DAsync(host, size_t(c->nx) * sizeof(double), c->tavg_acc[which], size_t(c->ld) * sizeof(double),
                             size_t(c->nx) * sizeof(double), size_t(c->ny) * (c->nz + 1), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if (total_time) *total_time = c->tavg_time;
    return 0;
}

int lesgo_gpu_tavg_reset(lesgo_gpu_ctx* c) {
    ENTER(c);
    if (!c) return 1;
    for (int i = 0; i < TA_N; ++i)
        if (c->tavg_acc[i]) CK(cudaMemsetAsync(c->tavg_acc[i], 0, size_t(c->plane) * (c->nz + 1) * sizeof(double), c->stream));
    c->tavg_time = 0.0;
    return 0;
}

int lesgo_gpu_turbines_init(lesgo_gpu_ctx* c, int nloc, const lesgo_gpu_turbine* t, int adm_correction) {
    ENTER(c);
    if (!c) return 1;
    return turbines_init(c, nloc, t, adm_correction);
}

int lesgo_gpu_turbines_rotation(lesgo_gpu_ctx* c, int nloc, const double* const* ind_t, const double* const* e_theta,
                                double tip_speed_ratio) {
    ENTER(c);
    if (!c) return 1;
    if (!c->turb_on) return c->fail("lesgo_gpu_turbines_rotation: call lesgo_gpu_turbines_init first");
    if (nloc != c->turb.nloc || (nloc > 0 && (!ind_t || !e_theta)))
        return c->fail("lesgo_gpu_turbines_rotation: nloc and the per-disk arrays must match lesgo_gpu_turbines_init");
    if (!(tip_speed_ratio != 0.0)) return c->fail("lesgo_gpu_turbines_rotation: tip_speed_ratio must be non-zero");
    std::vector<double> it, et;
    for (int s = 0; s < nloc; ++s) {
        const int n = c->turb_nodes[size_t(s) + 1] - c->turb_nodes[size_t(s)];
        if (n > 0 && (!ind_t[s] || !e_theta[s])) return c->fail("lesgo_gpu_turbines_rotation: disk without ind_t / e_theta");
        it.insert(it.end(), ind_t[s], ind_t[s] + n);
        et.insert(et.end(), e_theta[s], e_theta[s] + 3 * size_t(n));
    }
    if (upload_vec(c, &c->turb.ind_t, it, &c->turb_allocs) || upload_vec(c, &c->turb.e_theta, et, &c->turb_allocs)) return 1;
    c->turb.tip_speed_ratio = tip_speed_ratio;
    c->turb_fz = true;                     // e_theta has a z component: fza needs the interpolation to w nodes
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int lesgo_gpu_turbines_forcing(lesgo_gpu_ctx* c, double eps, double* u_d, double* u_d_T, double* f_n) {
    ENTER(c);
    if (!c) return 1;
    if (turbines_forcing(c, eps)) return 1;
    const size_t nb = size_t(c->turb.nloc) * sizeof(double);
    if (nb && (u_d || u_d_T || f_n)) {
        if (u_d) CK(cudaMemcpyAsync(u_d, c->turb.u_d, nb, cudaMemcpyDeviceToHost, c->stream));
        if (u_d_T) CK(cudaMemcpyAsync(u_d_T, c->turb.u_d_T, nb, cudaMemcpyDeviceToHost, c->stream));
        if (f_n) CK(cudaMemcpyAsync(f_n, c->turb.f_n, nb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// max(|u|)/dx, max(|v|)/dy, max(|w|)/dz over 1:nx, 1:ny, 1:nz-1 of this rank (cfl_util.f90:61-63,102-104)
static int local_inverse_dt(lesgo_gpu_ctx* c, double* m) {
    if (!c->red_dev) { if (dev_alloc(c, &c->red_dev, 8)) return 1; }
    if (!c->red_host) CK(cudaMallocHost(reinterpret_cast<void**>(&c->red_host), 8 * sizeof(double)));
    CK(cudaMemsetAsync(c->red_dev, 0, 8 * sizeof(double), c->stream));
    const int ids[3] = {LG_U, LG_V, LG_W};
    for (int i = 0; i < 3; ++i) {
        LG_LAUNCH(k_absmax, dim3(grid1d(long(c->nx / 2) * c->ny * (c->nz - 1))), dim3(kBlock), 0, c->stream,
                  field(c, ids[i]), c->lay(), c->nx, c->ny, 1, c->nz,
                  reinterpret_cast<unsigned long long*>(c->red_dev + i));
        c->launches++;
    }
    CK(cudaMemcpyAsync(c->red_host, c->red_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const double dx = c->d.L_x / c->nx, dy = c->d.L_y / c->ny;
    *m = std::fmax(c->red_host[0] / dx, std::fmax(c->red_host[1] / dy, c->red_host[2] / c->d.dz));
    return 0;
}

int lesgo_gpu_max_cfl(lesgo_gpu_ctx* c, double dt, double* cfl) {
    ENTER(c);
    // get_max_cfl, cfl_util.f90:35-69
    if (!c || !cfl) return 1;
    double m;
    if (local_inverse_dt(c, &m)) return 1;
    double r = dt * m;
    if (c->comm && c->comm->allreduce_max(&r, c->stream)) return c->fail(c->comm->error());
    *cfl = r;
    return 0;
}

int lesgo_gpu_cfl_dt(lesgo_gpu_ctx* c, double cfl, double* dt) {
    ENTER(c);
    // get_cfl_dt, cfl_util.f90:72-113
    if (!c || !dt) return 1;
    double m;
    if (local_inverse_dt(c, &m)) return 1;
    double r = cfl / m;
    if (c->comm && c->comm->allreduce(&r, 2, c->stream)) return c->fail(c->comm->error());
    *dt = r;
    return 0;
}

int lesgo_gpu_rmsdiv(lesgo_gpu_ctx* c, double* rms) {
    ENTER(c);
    // rmsdiv.f90:21-59 on the resident dudx, dvdy, dwdz
    if (!c || !rms) return 1;
    if (!c->red_dev) { if (dev_alloc(c, &c->red_dev, 8)) return 1; }
    if (!c->red_host) CK(cudaMallocHost(reinterpret_cast<void**>(&c->red_host), 8 * sizeof(double)));
    CK(cudaMemsetAsync(c->red_dev, 0, 8 * sizeof(double), c->stream));
    LG_LAUNCH(k_abs3sum, dim3(grid1d(long(c->nx / 2) * c->ny * (c->nz - 1))), dim3(kBlock), 0, c->stream,
              field(c, LG_DUDX), field(c, LG_DVDY), field(c, LG_DWDZ), c->lay(), c->nx, c->ny, 1, c->nz, c->red_dev + 4);
    c->launches++;
    CK(cudaMemcpyAsync(c->red_host, c->red_dev + 4, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double r = c->red_host[0] / (double(c->nx) * c->ny * (c->nz - 1));
    if (c->comm) {
        if (c->comm->allreduce_sum(&r, c->stream)) return c->fail(c->comm->error());
        r /= c->d.nproc;
    }
    *rms = r;
    return 0;
}

int lesgo_gpu_comm_unique_id(void* id128) { return lg::Comm::unique_id(id128, &g_err); }
int lesgo_gpu_comm_local_id(void* id128) { return lg::Comm::local_id(id128, &g_err); }

int lesgo_gpu_comm_init(lesgo_gpu_ctx* c, const void* id128) {
    ENTER(c);
    if (!c || !id128) return 1;
    if (c->comm) return c->fail("comm already initialised");
    std::string e;
    c->comm = lg::Comm::create(id128, c->d.coord, c->d.nproc, &e);
    if (!c->comm) return c->fail(e);
    return 0;
}

// ---- peer-memory transposes of the pressure solve -------------------------------------------------------
// Each rank owns one allocation of two pencil buffers (nproc blocks each, alternating between solves); the other
// ranks map it (same process: peer access; other process: CUDA IPC): the assembly kernel stores into it, the
// unpack kernel loads from it.
namespace {
struct P2PBlob {                 // 128 bytes, exchanged by the host like the NCCL id
    long long pid;
    int device, rank;
    unsigned long long ptr;
    unsigned long long doubles;  // size of one half
    unsigned char ipc[64];
    unsigned char pad[128 - 8 - 8 - 8 - 8 - 64];
};
static_assert(sizeof(P2PBlob) == 128, "blob is 128 bytes");
}  // namespace

int lesgo_gpu_comm_p2p_export(lesgo_gpu_ctx* c, void* blob128) {
    ENTER(c);
    if (!c || !blob128) return 1;
    if (c->d.nproc < 2 || c->d.nproc > 8) return c->fail("peer-memory transposes need 2..8 ranks on one node");
    const size_t half = size_t(c->nz) * ((c->ny + c->d.nproc - 1) / c->d.nproc) * c->ld * c->d.nproc;
    if (!c->p2p_buf) {
        if (dev_alloc(c, &c->p2p_buf, 2 * half + 16)) return 1;      // + 16 signal slots of the flag barrier (zeroed)
        if (dev_alloc(c, &c->p2p_flag, 1)) return 1;
        c->p2p_half = half;
        CK(cudaStreamSynchronize(c->stream));
    }
    P2PBlob b;
    std::memset(&b, 0, sizeof(b));
    b.pid = (long long)getpid(); b.device = c->device; b.rank = c->d.coord;
    b.ptr = (unsigned long long)(uintptr_t)c->p2p_buf; b.doubles = half;
#ifndef LESGO_EMUL
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->p2p_buf));
    static_assert(sizeof(h) <= sizeof(b.ipc), "ipc handle fits");
    std::memcpy(b.ipc, &h, sizeof(h));
#endif
    std::memcpy(blob128, &b, sizeof(b));
    return 0;
}

int lesgo_gpu_comm_p2p_import(lesgo_gpu_ctx* c, const void* blobs) {
    ENTER(c);
    if (!c) return 1;
    if (!blobs) { c->p2p_on = false; return 0; }                 // back to the NCCL all-to-alls
    if (!c->p2p_buf) return c->fail("lesgo_gpu_comm_p2p_export first");
    if (!c->comm) return c->fail("lesgo_gpu_comm_init first (the barrier of the peer-memory path rides on it)");
    const P2PBlob* b = static_cast<const P2PBlob*>(blobs);
    for (int q = 0; q < c->d.nproc; ++q) {
        if (b[q].rank != q || b[q].doubles != c->p2p_half) return c->fail("lesgo_gpu_comm_p2p_import: blobs out of order or other grid");
        double* base = nullptr;
        if (q == c->d.coord) base = c->p2p_buf;
        else if (b[q].pid == (long long)getpid()) {
            base = reinterpret_cast<double*>(uintptr_t(b[q].ptr));
#ifndef LESGO_EMUL
            if (b[q].device != c->device) {
                int ok = 0;
                CK(cudaDeviceCanAccessPeer(&ok, c->device, b[q].device));
                if (!ok) return c->fail("no peer access between the GPUs of ranks " + std::to_string(c->d.coord) + " and " + std::to_string(q));
                cudaError_t e = cudaDeviceEnablePeerAccess(b[q].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return c->fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
#endif
        } else {
#ifndef LESGO_EMUL
            cudaIpcMemHandle_t h;
            std::memcpy(&h, b[q].ipc, sizeof(h));
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return c->fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            base = static_cast<double*>(p);
            c->ipc_opened.push_back(p);
#else
            return c->fail("emulator: ranks must be threads of one process");
#endif
        }
        c->p2p_pencil[q] = base;
        c->p2p_alt[q] = base + c->p2p_half;
        c->p2p_sig[q] = reinterpret_cast<unsigned long long*>(base + 2 * c->p2p_half);
    }
    c->p2p_on = true;
    {
        // LESGO_P2P_FLAGS=0: keep the one-double NCCL all-reduce as the barrier of the peer-memory transposes
        const char* e = std::getenv("LESGO_P2P_FLAGS");
        c->p2p_flags_on = !(e && e[0] == '0');
#ifndef LESGO_EMUL
        // The flag barrier is a kernel that spins until every peer's barrier kernel has run, which needs the peers'
        // kernels to be able to run WHILE it spins.  That holds with one rank per GPU.  Ranks that share a device
        // (the single-device transport of the tests) also share its few hardware launch queues, where a peer's
        // pending kernels can sit behind the spinning one: there the host-ordered barrier of the transport is used.
        // (decided from the blobs every rank holds, so all ranks decide alike)
        for (int q = 0; q < c->d.nproc; ++q)
            for (int r = q + 1; r < c->d.nproc; ++r)
                if (b[q].pid == b[r].pid && b[q].device == b[r].device) c->p2p_flags_on = false;   // threads of one process on one GPU
#endif
    }
    return 0;
}

int lesgo_gpu_sync_real_array(lesgo_gpu_ctx* c, double* var, int isync) {
    ENTER(c);
    if (!c || !var) return 1;
    if (c->d.nproc == 1) return 0;
    if (!c->comm) return c->fail("lesgo_gpu_comm_init has not been called");
    Staged st(c);
    double* d = st.in(var, NFIELD, true, true);
    if (!d) return 1;
    if (c->comm->sync_planes(d, c->plane, c->nz, isync, c->stream)) return c->fail(c->comm->error());
    return st.finish();
}

int lesgo_gpu_padd(lesgo_gpu_ctx* c, double* u_big, const double* u, int nplanes) {
    ENTER(c);
    if (!c || !u_big || !u || nplanes < 1) return 1;
    Staged st(c);
    double* du = st.in(u, size_t(c->plane) * nplanes, true, false);
    double* db = st.in(u_big, size_t(c->plane_big) * nplanes, false, true);
    if (!du || !db) return 1;
    LG_LAUNCH(k_padd, dim3(grid1d(long(c->lh_big) * c->ny2 * nplanes)), dim3(kBlock), 0, c->stream, du, db, c->nx,
              c->ny, c->ld, c->ny2, c->ld_big, nplanes);
    c->launches++;
    return st.finish();
}

int lesgo_gpu_unpadd(lesgo_gpu_ctx* c, double* cc, const double* cc_big, int nplanes) {
    ENTER(c);
    if (!c || !cc || !cc_big || nplanes < 1) return 1;
    Staged st(c);
    double* db = st.in(cc_big, size_t(c->plane_big) * nplanes, true, false);
    double* dc = st.in(cc, size_t(c->plane) * nplanes, false, true);
    if (!dc || !db) return 1;
    LG_LAUNCH(k_unpadd, dim3(grid1d(long(c->lh) * c->ny * nplanes)), dim3(kBlock), 0, c->stream, dc, db, c->nx,
              c->ny, c->ld, c->ny2, c->ld_big, nplanes);
    c->launches++;
    return st.finish();
}

int lesgo_gpu_test_filter(lesgo_gpu_ctx* c, double* f, const double* G, int nplanes) {
    ENTER(c);
    // test_filtermodule.f90:126-146: r2c, multiply by the real kernel G(lh, ny), c2r
    if (!c || !f || !G || nplanes < 1) return 1;
    if (nplanes > c->nz + 1) return c->fail("test_filter: at most nz+1 planes per call");
    if (need_small(c, 2)) return 1;
    Staged st(c);
    double* df = st.in(f, size_t(c->plane) * nplanes, true, true);
    double* dg = st.in(G, size_t(c->lh) * c->ny, true, false);
    if (!df || !dg) return 1;
    ProScale ps; ps.src[0] = df; ps.lay = c->lay(); ps.scale = 1.0;
    double* d0[1] = {c->sa[0]};
    if (xfwd(c, false, ps, 1, d0, c->plane, c->ld, c->nx / 2, c->ny, 0, nplanes)) return 1;
    YArgs a = yargs(c, c->plane, c->ld, c->plane, c->ld, c->nx / 2, 0);
    a.fld[0].src = c->sa[0]; a.fld[0].out[0] = YOutSpec{c->sa[1], Y_TABLE};
    a.table = dg; a.table_row = c->lh;
    if (ypass(c, c->ny, c->ny, a, 1, 0, nplanes)) return 1;
    const double* s0[1] = {c->sa[1]};
    double* o0[1] = {df};
    if (xinv(c, false, s0, c->plane, c->ld, c->nx / 2, 1, o0, c->lay(), c->ny, 0, nplanes)) return 1;
    return st.finish();
}

int lesgo_gpu_tridag_array(lesgo_gpu_ctx* c, const double* a, const double* b, const double* cc, const double* r,
                           double* u, int n) {
    ENTER(c);
    if (!c || !a || !b || !cc || !r || !u || n < 2) return 1;
    // the serial form only (tridag_array.f90:166-246): the MPI form pipelines ONE system across the ranks, which
    // this library replaces inside press_stag_array by the slab <-> pencil transposes
    if (c->d.nproc > 1)
        return c->fail("tridag_array: general-coefficient entry point is single-slab only; with nproc > 1 the "
                       "distributed solve lives inside lesgo_gpu_press_stag_array");
    Staged st(c);
    const size_t nc = size_t(c->lh) * c->ny * n, nr = size_t(c->plane) * n;
    double* da = st.in(a, nc, true, false);
    double* db = st.in(b, nc, true, false);
    double* dc = st.in(cc, nc, true, false);
    double* dr = st.in(r, nr, true, false);
    double* du = st.in(u, nr, true, true);
    if (!da || !db || !dc || !dr || !du) return 1;
    void* work = nullptr;
    CK(cudaMalloc(&work, nc * sizeof(double) + 16));
    struct Free { void* p; lesgo_gpu_ctx* c; ~Free() { cudaStreamSynchronize(c->stream); cudaFree(p); } } free_work{work, c};
    int* flag = reinterpret_cast<int*>(static_cast<double*>(work) + nc);
    CK(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
    const int nm = (c->lh - 1) * c->ny;
    LG_LAUNCH(k_tridag_general, dim3((nm + 127) / 128), dim3(128), 0, c->stream, c->lh, c->ny, n, c->ld, da, db, dc,
              dr, du, static_cast<double*>(work), flag);
    c->launches++;
    int hflag = 0;
    CK(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    int rc = st.finish();
    CK(cudaStreamSynchronize(c->stream));
    if (rc) return rc;
    if (hflag) return c->fail("tridag_array failed: zero pivot (tridag_array.f90:55-60,101-108)");
    return 0;
}

}  // extern "C"
