// comm.h -- the communication backend that replaces mpi_defs.f90's MPI point-to-point
// calls: NCCL send/recv over NVLink between the z-slab ranks (one rank = one GPU).
// NCCL is loaded with dlopen so the library has no link-time dependency on it and a
// single-GPU run never touches it.
#pragma once
#include <string>

#include "portable.h"

namespace lg {

class Comm {
public:
    // 128-byte ncclUniqueId for the host to broadcast (MPI_Bcast / torch.distributed)
    static int unique_id(void* id128, std::string* err);
    // id of the single-device transport: all ranks are threads of one process sharing one GPU (comm.cu)
    static int local_id(void* id128, std::string* err);
    static Comm* create(const void* id128, int rank, int nranks, std::string* err);
    virtual ~Comm() {}
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }
    const std::string& error() const { return err_; }

    // One fused exchange: for every i, send sendbuf[i] (count doubles) to `dest[i]` and
    // receive recvbuf[i] from `src[i]`; ranks outside [0, nranks) are MPI_PROC_NULL
    // (mpi_defs.f90:79-83): nothing is sent, the receive buffer is left untouched.
    virtual int exchange(int n, const double* const* sendbuf, const int* dest, double* const* recvbuf,
                         const int* src, const size_t* count, cudaStream_t s) = 0;
    // host scalar all-reduce (cfl_util.f90:66,107; rmsdiv.f90:54)
    virtual int allreduce(double* host_value, int op /*0 sum, 1 max, 2 min*/, cudaStream_t s) = 0;
    // in-place sum of n doubles in DEVICE memory over all ranks, stream-ordered, no host sync
    // (turbines.f90:553-560: MPI_Allreduce of the per-disk velocities)
    virtual int allreduce_sum_dev(double* dev, size_t n, cudaStream_t s) = 0;
    // every rank r sends sendbuf + r*count and receives into recvbuf + r*count (transpose)
    virtual int alltoall(const double* sendbuf, double* recvbuf, size_t count, cudaStream_t s) = 0;

    // mpi_sync_real_array (mpi_defs.f90:167-264): isync bit 0 = SYNC_DOWN (k=1 of coord+1 ->
    // k=nz of coord), bit 1 = SYNC_UP (k=nz-1 of coord -> k=0 of coord+1)
    int sync_planes(double* var, long plane, int nz, int isync, cudaStream_t s) {
        const double* sb[2];
        double* rb[2];
        int dest[2], src[2];
        size_t cnt[2];
        int n = 0;
        if (isync & 1) { sb[n] = var + plane; dest[n] = rank_ - 1; rb[n] = var + plane * nz; src[n] = rank_ + 1; cnt[n] = size_t(plane); ++n; }
        if (isync & 2) { sb[n] = var + plane * (nz - 1); dest[n] = rank_ + 1; rb[n] = var; src[n] = rank_ - 1; cnt[n] = size_t(plane); ++n; }
        return exchange(n, sb, dest, rb, src, cnt, s);
    }
    int allreduce_max(double* v, cudaStream_t s) { return allreduce(v, 1, s); }
    int allreduce_sum(double* v, cudaStream_t s) { return allreduce(v, 0, s); }

protected:
    int rank_ = 0, nranks_ = 1;
    std::string err_;
};

}  // namespace lg
