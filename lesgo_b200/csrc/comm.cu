// comm.cu -- NCCL backend of lg::Comm (dlopen'ed libnccl.so.2).
#include "comm.h"

#ifdef LESGO_EMUL
// Emulator build: ranks are threads of one process; messages go through in-process queues.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <vector>
namespace lg {
namespace {
struct World {
    std::mutex m;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<std::vector<double>>> q;   // (src, dst) -> FIFO
};
std::mutex g_worlds_m;
std::map<std::string, std::shared_ptr<World>> g_worlds;

class EmulComm : public Comm {
public:
    std::shared_ptr<World> w;
    void put(int dst, const double* p, size_t n) {
        std::lock_guard<std::mutex> l(w->m);
        w->q[{rank_, dst}].emplace_back(p, p + n);
        w->cv.notify_all();
    }
    void get(int src, double* p, size_t n) {
        std::unique_lock<std::mutex> l(w->m);
        auto& dq = w->q[{src, rank_}];
        w->cv.wait(l, [&] { return !dq.empty(); });
        std::memcpy(p, dq.front().data(), n * sizeof(double));
        dq.pop_front();
    }
    int exchange(int n, const double* const* sendbuf, const int* dest, double* const* recvbuf, const int* src,
                 const size_t* count, cudaStream_t) override {
        for (int i = 0; i < n; ++i) if (dest[i] >= 0 && dest[i] < nranks_) put(dest[i], sendbuf[i], count[i]);
        for (int i = 0; i < n; ++i) if (src[i] >= 0 && src[i] < nranks_) get(src[i], recvbuf[i], count[i]);
        return 0;
    }
    int allreduce(double* v, int op, cudaStream_t) override {
        for (int r = 0; r < nranks_; ++r) if (r != rank_) put(r, v, 1);
        double acc = *v;
        for (int r = 0; r < nranks_; ++r) {
            if (r == rank_) continue;
            double x; get(r, &x, 1);
            acc = op == 0 ? acc + x : (op == 1 ? (x > acc ? x : acc) : (x < acc ? x : acc));
        }
        *v = acc;
        return 0;
    }
    int allreduce_sum_dev(double* dev, size_t n, cudaStream_t) override {
        for (int r = 0; r < nranks_; ++r) if (r != rank_) put(r, dev, n);
        std::vector<double> x(n);
        // rank order, so every rank forms the same sum
        std::vector<double> own(dev, dev + n), acc(n, 0.0);
        for (int r = 0; r < nranks_; ++r) {
            const double* src = own.data();
            if (r != rank_) { get(r, x.data(), n); src = x.data(); }
            for (size_t i = 0; i < n; ++i) acc[i] = r == 0 ? src[i] : acc[i] + src[i];
        }
        std::memcpy(dev, acc.data(), n * sizeof(double));
        return 0;
    }
    int alltoall(const double* sendbuf, double* recvbuf, size_t count, cudaStream_t) override {
        for (int r = 0; r < nranks_; ++r) put(r, sendbuf + size_t(r) * count, count);
        for (int r = 0; r < nranks_; ++r) get(r, recvbuf + size_t(r) * count, count);
        return 0;
    }
    void set(int r, int n) { rank_ = r; nranks_ = n; }
};
}  // namespace
int Comm::unique_id(void* id128, std::string*) {
    std::random_device rd;
    unsigned char* b = static_cast<unsigned char*>(id128);
    for (int i = 0; i < 128; ++i) b[i] = static_cast<unsigned char>('a' + rd() % 26);
    return 0;
}
int Comm::local_id(void* id128, std::string* e) { return unique_id(id128, e); }
Comm* Comm::create(const void* id128, int rank, int nranks, std::string*) {
    std::string key(static_cast<const char*>(id128), 128);
    EmulComm* c = new EmulComm;
    c->set(rank, nranks);
    std::lock_guard<std::mutex> l(g_worlds_m);
    auto& w = g_worlds[key];
    if (!w) w = std::make_shared<World>();
    c->w = w;
    return c;
}
}  // namespace lg
#else
#include <dlfcn.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <vector>

namespace lg {
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };                 // nccl.h: ncclDouble = 8
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Api {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(f, name) f = reinterpret_cast<decltype(f)>(dlsym(h, name)); if (!f) { err = std::string("dlsym ") + name; return false; }
        SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
        SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce")
        SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
Api& api() { static Api a; return a; }

class NcclComm : public Comm {
public:
    ncclComm_t comm = nullptr;
    double* dbuf = nullptr;     // device scalar for all-reduce
    double* hbuf = nullptr;
    ~NcclComm() override {
        if (comm) api().CommDestroy(comm);
        if (dbuf) cudaFree(dbuf);
        if (hbuf) cudaFreeHost(hbuf);
    }
    int fail(const char* what, int rc) { err_ = std::string(what) + ": " + api().GetErrorString(rc); return 1; }
    int exchange(int n, const double* const* sendbuf, const int* dest, double* const* recvbuf, const int* src,
                 const size_t* count, cudaStream_t s) override {
        int rc = api().GroupStart();
        if (rc) return fail("ncclGroupStart", rc);
        for (int i = 0; i < n; ++i) {
            if (dest[i] >= 0 && dest[i] < nranks_) {
                rc = api().Send(sendbuf[i], count[i], ncclFloat64, dest[i], comm, s);
                if (rc) return fail("ncclSend", rc);
            }
            if (src[i] >= 0 && src[i] < nranks_) {
                rc = api().Recv(recvbuf[i], count[i], ncclFloat64, src[i], comm, s);
                if (rc) return fail("ncclRecv", rc);
            }
        }
        rc = api().GroupEnd();
        if (rc) return fail("ncclGroupEnd", rc);
        return 0;
    }
    int allreduce(double* v, int op, cudaStream_t s) override {
        if (!dbuf) { cudaMalloc(reinterpret_cast<void**>(&dbuf), sizeof(double)); cudaMallocHost(reinterpret_cast<void**>(&hbuf), sizeof(double)); }
        *hbuf = *v;
        cudaMemcpyAsync(dbuf, hbuf, sizeof(double), cudaMemcpyHostToDevice, s);
        int rc = api().AllReduce(dbuf, dbuf, 1, ncclFloat64, op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin), comm, s);
        if (rc) return fail("ncclAllReduce", rc);
        cudaMemcpyAsync(hbuf, dbuf, sizeof(double), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { err_ = "allreduce sync failed"; return 1; }
        *v = *hbuf;
        return 0;
    }
    int allreduce_sum_dev(double* dev, size_t n, cudaStream_t s) override {
        int rc = api().AllReduce(dev, dev, n, ncclFloat64, ncclSum, comm, s);
        if (rc) return fail("ncclAllReduce", rc);
        return 0;
    }
    int alltoall(const double* sendbuf, double* recvbuf, size_t count, cudaStream_t s) override {
        int rc = api().GroupStart();
        if (rc) return fail("ncclGroupStart", rc);
        for (int r = 0; r < nranks_; ++r) {
            rc = api().Send(sendbuf + size_t(r) * count, count, ncclFloat64, r, comm, s);
            if (rc) return fail("ncclSend", rc);
            rc = api().Recv(recvbuf + size_t(r) * count, count, ncclFloat64, r, comm, s);
            if (rc) return fail("ncclRecv", rc);
        }
        rc = api().GroupEnd();
        if (rc) return fail("ncclGroupEnd", rc);
        return 0;
    }
    friend class Comm;
    void set(int r, int n) { rank_ = r; nranks_ = n; }
};
// ---- single-device transport ---------------------------------------------------------------------------------
// The z-slab ranks are threads of ONE process that share ONE GPU (every context on the same device, each with its
// own stream): the multi-slab path -- ghost-plane halos, slab <-> pencil transposes of the pressure solve, the
// k = 0 chain, disk-velocity reductions -- then runs on a single B200, which is how a one-GPU box exercises it
// (tests/test_gpu_parity.py).  Messages are device-to-device copies ordered by CUDA events; NCCL cannot do this
// (it refuses two ranks on one device).  Selected by an id made with Comm::local_id.
const char kLocalTag[] = "LESGO-LOCAL-COMM:";
struct LWorld {
    struct Msg { double* buf; size_t n; cudaEvent_t ev; };
    std::mutex m;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<Msg>> q;                   // (src, dst) -> device messages
    std::map<std::pair<int, int>, std::deque<std::vector<double>>> hq;  // (src, dst) -> host scalars
};
std::mutex g_lworlds_m;
std::map<std::string, std::shared_ptr<LWorld>> g_lworlds;

class LocalComm : public Comm {
public:
    std::shared_ptr<LWorld> w;
    void set(int r, int n) { rank_ = r; nranks_ = n; }
    int put(int dst, const double* p, size_t n, cudaStream_t s) {
        // stream-ordered allocation: no call on this path may synchronise the DEVICE, because another rank's
        // flag-barrier kernel may be spinning on it until this rank gets to launch its own (ops.h k_p2p_barrier)
        LWorld::Msg m{nullptr, n, nullptr};
        if (cudaMallocAsync(reinterpret_cast<void**>(&m.buf), (n ? n : 1) * sizeof(double), s) != cudaSuccess) { err_ = "local comm: cudaMallocAsync"; return 1; }
        cudaMemcpyAsync(m.buf, p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
        cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming);
        cudaEventRecord(m.ev, s);
        std::lock_guard<std::mutex> l(w->m);
        w->q[{rank_, dst}].push_back(m);
        w->cv.notify_all();
        return 0;
    }
    int get(int src, double* p, size_t n, cudaStream_t s) {
        LWorld::Msg m;
        {
            std::unique_lock<std::mutex> l(w->m);
            auto& dq = w->q[{src, rank_}];
            w->cv.wait(l, [&] { return !dq.empty(); });
            m = dq.front();
            dq.pop_front();
        }
        if (m.n != n) { err_ = "local comm: message size mismatch"; return 1; }
        cudaStreamWaitEvent(s, m.ev, 0);
        cudaMemcpyAsync(p, m.buf, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
        const cudaError_t e = cudaFreeAsync(m.buf, s);          // after the copy, in stream order
        cudaEventDestroy(m.ev);
        if (e != cudaSuccess) { err_ = std::string("local comm: ") + cudaGetErrorString(e); return 1; }
        return 0;
    }
    void hput(int dst, const std::vector<double>& v) {
        std::lock_guard<std::mutex> l(w->m);
        w->hq[{rank_, dst}].push_back(v);
        w->cv.notify_all();
    }
    std::vector<double> hget(int src) {
        std::unique_lock<std::mutex> l(w->m);
        auto& dq = w->hq[{src, rank_}];
        w->cv.wait(l, [&] { return !dq.empty(); });
        std::vector<double> v = dq.front();
        dq.pop_front();
        return v;
    }
    int exchange(int n, const double* const* sendbuf, const int* dest, double* const* recvbuf, const int* src,
                 const size_t* count, cudaStream_t s) override {
        for (int i = 0; i < n; ++i) if (dest[i] >= 0 && dest[i] < nranks_) if (put(dest[i], sendbuf[i], count[i], s)) return 1;
        for (int i = 0; i < n; ++i) if (src[i] >= 0 && src[i] < nranks_) if (get(src[i], recvbuf[i], count[i], s)) return 1;
        return 0;
    }
    // the host forms the sum in rank order, so every rank holds the same bits
    int reduce_host(std::vector<double>& v, int op) {
        for (int r = 0; r < nranks_; ++r) if (r != rank_) hput(r, v);
        std::vector<double> acc;
        for (int r = 0; r < nranks_; ++r) {
            const std::vector<double> x = r == rank_ ? v : hget(r);
            if (r == 0) { acc = x; continue; }
            for (size_t i = 0; i < x.size(); ++i)
                acc[i] = op == 0 ? acc[i] + x[i] : (op == 1 ? (x[i] > acc[i] ? x[i] : acc[i]) : (x[i] < acc[i] ? x[i] : acc[i]));
        }
        v = acc;
        return 0;
    }
    int allreduce(double* v, int op, cudaStream_t s) override {
        if (cudaStreamSynchronize(s) != cudaSuccess) { err_ = "allreduce sync failed"; return 1; }
        std::vector<double> x(1, *v);
        reduce_host(x, op);
        *v = x[0];
        return 0;
    }
    int allreduce_sum_dev(double* dev, size_t n, cudaStream_t s) override {
        std::vector<double> x(n);
        cudaMemcpyAsync(x.data(), dev, n * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { err_ = "allreduce sync failed"; return 1; }
        reduce_host(x, 0);             // every rank's earlier work on its stream is complete once this returns
        cudaMemcpyAsync(dev, x.data(), n * sizeof(double), cudaMemcpyHostToDevice, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { err_ = "allreduce sync failed"; return 1; }
        return 0;
    }
    int alltoall(const double* sendbuf, double* recvbuf, size_t count, cudaStream_t s) override {
        for (int r = 0; r < nranks_; ++r) if (put(r, sendbuf + size_t(r) * count, count, s)) return 1;
        for (int r = 0; r < nranks_; ++r) if (get(r, recvbuf + size_t(r) * count, count, s)) return 1;
        return 0;
    }
};
}  // namespace

int Comm::local_id(void* id128, std::string*) {
    std::random_device rd;
    unsigned char* b = static_cast<unsigned char*>(id128);
    std::memset(b, 0, 128);
    std::memcpy(b, kLocalTag, sizeof(kLocalTag) - 1);
    for (int i = int(sizeof(kLocalTag)) - 1; i < 127; ++i) b[i] = static_cast<unsigned char>('a' + rd() % 26);
    return 0;
}

int Comm::unique_id(void* id128, std::string* err) {
    if (!api().load()) { if (err) *err = api().err; return 1; }
    ncclUniqueId id;
    int rc = api().GetUniqueId(&id);
    if (rc) { if (err) *err = std::string("ncclGetUniqueId: ") + api().GetErrorString(rc); return 1; }
    std::memcpy(id128, &id, 128);
    return 0;
}

Comm* Comm::create(const void* id128, int rank, int nranks, std::string* err) {
    if (std::memcmp(id128, kLocalTag, sizeof(kLocalTag) - 1) == 0) {
        LocalComm* c = new LocalComm;
        c->set(rank, nranks);
        std::lock_guard<std::mutex> l(g_lworlds_m);
        auto& w = g_lworlds[std::string(static_cast<const char*>(id128), 128)];
        if (!w) w = std::make_shared<LWorld>();
        c->w = w;
        return c;
    }
    if (!api().load()) { if (err) *err = api().err; return nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    NcclComm* c = new NcclComm;
    c->set(rank, nranks);
    int rc = api().CommInitRank(&c->comm, nranks, id, rank);
    if (rc) { if (err) *err = std::string("ncclCommInitRank: ") + api().GetErrorString(rc); delete c; return nullptr; }
    return c;
}

}  // namespace lg
#endif
