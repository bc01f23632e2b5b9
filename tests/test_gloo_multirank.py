"""world_size-2 gloo tests (CPU) of the N>1 host logic: slab decomposition helpers, the comm
plumbing, and the unique-id bootstrap.  The arithmetic run over gloo is the ORACLE's (checker),
driven through lesgo_b200.slab's decomposition + TorchComm, and must equal the single-slab run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lesgo_b200 import slab
from oracle import lesgo_oracle as O

KW = dict(nx=16, ny=12, Nz=8, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False,
          use_mean_p_force=True, mean_p_force_x=1.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeCore:
    """Stands in for lesgo_b200.Core (which needs a GPU): records the id it is initialised with."""

    def __init__(self, rank):
        self.rank, self.got = rank, None

    def comm_unique_id(self):
        return bytes((7 * i + 3) % 256 for i in range(128))

    def comm_init(self, ident):
        self.got = ident

    # peer-memory bootstrap: every rank must receive all exports, in rank order
    def comm_p2p_export(self):
        return bytes([self.rank]) * 128

    def comm_p2p_import(self, blobs):
        if blobs is not None:
            self.blobs = blobs


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) bootstrap: every rank ends up with rank 0's id
        core = FakeCore(rank)
        ident = slab.bootstrap_comm(core, dist)
        assert core.got == ident == bytes((7 * i + 3) % 256 for i in range(128))
        assert core.blobs == [bytes([r]) * 128 for r in range(world)]
        # (2) decomposition + comm: two oracle steps over gloo
        pg = O.Params(nproc=1, **KW)
        ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=world, seed=3, amp=0.3)
        p = O.Params(nproc=world, coord=rank, **KW)
        assert p.nz == slab.local_nz(KW["Nz"], world) and p.nz_tot == slab.nz_total(KW["Nz"], world)
        sp = O.Spectral(p)
        G = O.test_filter_kernel(sp)
        s = O.State(p)
        s.u, s.v, s.w = (slab.scatter_slab(f, rank, world) for f in (ug, vg, wg))
        assert np.array_equal(s.u, O.scatter_slab(ug, p))
        comm = slab.TorchComm(dist)
        for it in range(2):
            O.step(s, sp, comm, mode="full", first_step=(it == 0), G_test=G)
        cfl = O.get_max_cfl(s, p, comm)
        np.savez(os.path.join(tmp, f"rank{rank}.npz"), u=s.u, w=s.w, p=s.p, RHSx=s.RHSx, cfl=cfl)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_slab(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    pg = O.Params(nproc=1, **KW)
    spg = O.Spectral(pg)
    Gg = O.test_filter_kernel(spg)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=world, seed=3, amp=0.3)
    s = O.State(pg)
    s.u, s.v, s.w = (O.scatter_slab(f, pg) for f in (ug, vg, wg))
    for it in range(2):
        O.step(s, spg, O.LocalComm(), mode="full", first_step=(it == 0), G_test=Gg)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    nzt = pg.nz_tot
    for n, top in (("u", False), ("w", True), ("p", True), ("RHSx", False)):
        g = slab.gather_slabs([r[n] for r in res], world, top_plane=top)
        hi = nzt if top else nzt - 1
        ref = getattr(s, n)[1:hi + 1, :, :pg.nx]
        err = np.linalg.norm((g[1:hi + 1, :, :pg.nx] - ref).ravel()) / np.linalg.norm(ref.ravel())
        assert err < 1e-12, (n, err)
    cfl_ref = O.get_max_cfl(s, pg, O.LocalComm())
    assert all(abs(float(r["cfl"]) - cfl_ref) < 1e-14 for r in res)


def test_decomposition_helpers():
    assert slab.local_nz(256, 8) == 33 and slab.nz_total(256, 8) == 257
    assert slab.local_nz(64, 4) == 17 and slab.nz_total(64, 4) == 65
    g = np.arange(10 * 2 * 4, dtype=float).reshape(10, 2, 4)                  # levels 1..9
    parts = [slab.scatter_slab(g, c, 4) for c in range(4)]                    # nz = 3
    assert parts[0].shape[0] == 4 and np.all(parts[0][0] == slab.BOGUS)
    assert np.array_equal(parts[1][0], parts[0][2]) and np.array_equal(parts[0][3], parts[1][1])
    back = slab.gather_slabs(parts, 4, top_plane=True)
    assert np.array_equal(back[1:], g[1:])
