// comm.cu -- NCCL backend of lg::Comm (dlopen'ed libnccl.so.2).
#include "comm.h"

#ifdef LESGO_EMUL
namespace lg {
int Comm::unique_id(void* id128, std::string* err) { (void)id128; if (err) *err = "emulator build has no NCCL"; return 1; }
Comm* Comm::create(const void*, int, int, std::string* err) { if (err) *err = "emulator build has no NCCL"; return nullptr; }
}  // namespace lg
#else
#include <dlfcn.h>

#include <cstring>
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
}  // namespace

int Comm::unique_id(void* id128, std::string* err) {
    if (!api().load()) { if (err) *err = api().err; return 1; }
    ncclUniqueId id;
    int rc = api().GetUniqueId(&id);
    if (rc) { if (err) *err = std::string("ncclGetUniqueId: ") + api().GetErrorString(rc); return 1; }
    std::memcpy(id128, &id, 128);
    return 0;
}

Comm* Comm::create(const void* id128, int rank, int nranks, std::string* err) {
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
