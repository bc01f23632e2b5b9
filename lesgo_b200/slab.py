"""z-slab decomposition helpers and the multi-rank bootstrap (host side).

LESGO's only parallelism is a 1-D slab decomposition in z (mpi_defs.f90:77-87,
input_util.f90:197-201): rank `coord` owns levels 1..nz-1 of its `(ld, ny, 0:nz)` arrays,
plane 0 mirrors plane nz-1 of the rank below and plane nz mirrors plane 1 of the rank above
(mpi_defs.f90:177-180).  One rank = one process = one GPU; `torch.distributed` is only the
plumbing that carries the NCCL unique id to every rank (the Fortran shim uses MPI_Bcast for
the same thing) -- the data path talks NCCL inside liblesgo_cuda.so.
"""
from __future__ import annotations

import numpy as np

BOGUS = -1234567890.0   # param.f90:93


def local_nz(Nz: int, nproc: int) -> int:
    """Per-rank nz (input_util.f90:197): each rank holds planes 0..nz."""
    return Nz // nproc + 1


def nz_total(Nz: int, nproc: int) -> int:
    """nz_tot after LESGO re-derives it (input_util.f90:200)."""
    return (local_nz(Nz, nproc) - 1) * nproc + 1


def global_level(coord: int, k: int, nz: int) -> int:
    """Global w/uv level of local plane k on rank `coord` (grid.f90:82)."""
    return coord * (nz - 1) + k


def scatter_slab(g: np.ndarray, coord: int, nproc: int) -> np.ndarray:
    """Global array with levels 1..nz_tot at index 1.. -> this rank's (nz+1, ny, ld) slab with
    ghost planes filled from the neighbours' owned planes; levels outside the domain are BOGUS
    (initial.f90:178-182)."""
    nzt = g.shape[0] - 1
    nz = (nzt - 1) // nproc + 1
    loc = np.full((nz + 1,) + g.shape[1:], BOGUS)
    for k in range(nz + 1):
        gk = global_level(coord, k, nz)
        if 1 <= gk <= nzt:
            loc[k] = g[gk]
    return loc


def gather_slabs(locs, nproc: int, top_plane: bool = False) -> np.ndarray:
    """Owned planes (1..nz-1, plus nz of the top rank when top_plane) of every rank -> global."""
    nz = locs[0].shape[0] - 1
    nzt = (nz - 1) * nproc + 1
    g = np.full((nzt + 1,) + locs[0].shape[1:], np.nan)
    for coord, loc in enumerate(locs):
        hi = nz if (coord == nproc - 1 and top_plane) else nz - 1
        for k in range(1, hi + 1):
            g[global_level(coord, k, nz)] = loc[k]
    return g


class TorchComm:
    """MPI-like blocking send/recv/sendrecv/allreduce on numpy buffers (out-of-range ranks =
    MPI_PROC_NULL, mpi_defs.f90:79-83) over torch.distributed; with gloo on CPU it drives the
    world_size-2 tests of the decomposition logic."""

    def __init__(self, dist):
        self.dist = dist
        self.coord = dist.get_rank()
        self.nproc = dist.get_world_size()

    def _ok(self, r):
        return 0 <= r < self.nproc

    def send(self, buf, dest, tag):
        if self._ok(dest):
            import torch
            self.dist.send(torch.from_numpy(np.ascontiguousarray(buf, dtype=np.float64).copy()), dst=dest)

    def recv(self, buf, src, tag):
        if self._ok(src):
            import torch
            t = torch.empty(buf.shape, dtype=torch.float64)
            self.dist.recv(t, src=src)
            buf[...] = t.numpy()

    def sendrecv(self, sendbuf, dest, recvbuf, src, tag):
        import torch
        ops = []
        rt = None
        if self._ok(dest):
            ops.append(self.dist.P2POp(self.dist.isend, torch.from_numpy(
                np.ascontiguousarray(sendbuf, dtype=np.float64).copy()), dest))
        if self._ok(src):
            rt = torch.empty(recvbuf.shape, dtype=torch.float64)
            ops.append(self.dist.P2POp(self.dist.irecv, rt, src))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        if rt is not None:
            recvbuf[...] = rt.numpy()

    def allreduce(self, value, op):
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64)
        o = {"min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}[op]
        self.dist.all_reduce(t, op=o)
        return float(t.item())


def bootstrap_comm(core, dist) -> bytes:
    """Make rank 0's NCCL unique id known to every rank and initialise the library's communicator
    (the analogue of the MPI_Bcast in fortran/lesgo_gpu_mod.f90).  `core` needs `comm_unique_id()`
    and `comm_init(bytes)`; `dist` is an initialised torch.distributed."""
    ident = [core.comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    if not isinstance(ident[0], (bytes, bytearray)) or len(ident[0]) != 128:
        raise RuntimeError("unique id broadcast failed")
    core.comm_init(bytes(ident[0]))
    # peer-memory pressure transposes (LESGO_P2P=0: keep the NCCL all-to-alls)
    import os
    if os.environ.get("LESGO_P2P", "1") != "0" and 2 <= dist.get_world_size() <= 8:
        # collective and all-or-nothing: a rank that cannot export or map a peer buffer (no peer access,
        # IPC refused) makes EVERY rank stay on the NCCL path, otherwise the ranks would disagree and hang
        world = dist.get_world_size()
        try:
            mine, err = core.comm_p2p_export(), None
        except Exception as e:  # noqa
            mine, err = None, str(e)
        blobs = [None] * world
        dist.all_gather_object(blobs, mine)
        ok = all(b is not None for b in blobs)
        if ok:
            try:
                core.comm_p2p_import([bytes(b) for b in blobs])
            except Exception as e:  # noqa
                ok, err = False, str(e)
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if not all(oks):
            core.comm_p2p_import(None)
            if dist.get_rank() == 0:
                import sys
                print(f"[lesgo_b200] peer-memory transposes unavailable ({err or 'another rank failed'}): "
                      "using NCCL all-to-alls", file=sys.stderr)
    return bytes(ident[0])
