"""Multi-GPU parity: the z-slab ranks (one lesgo_b200.Core per GPU, here as threads of one
process, NCCL between them) must reproduce the single-slab oracle.  Needs >= 2 GPUs."""
import pytest

import lesgo_b200
from helpers import check_multirank_steps

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_steps(nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, nproc, nsteps=2, tol=1e-11,
                                device_of=lambda coord: coord)
    print(nproc, out)


def test_multi_gpu_full_step():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, 2, nsteps=2, tol=1e-11, mode="full",
                                device_of=lambda coord: coord)
    print(out)


def test_multi_gpu_lasd():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, dt=2e-3)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, 2, nsteps=4, tol=1e-11, mode="full", lasd=True,
                                device_of=lambda coord: coord)
    print(out)


def test_multi_gpu_turbines():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, 2, nsteps=2, tol=1e-11, mode="full", turbines=True, tavg=True,
                                device_of=lambda coord: coord)
    print(out)


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_p2p_transposes(nproc):
    """The pressure transposes over NVLink peer memory must give the same bits as the NCCL all-to-alls
    (both reproduce the single-slab oracle to 1e-11)."""
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, nproc, nsteps=3, tol=1e-11, p2p=True,
                                device_of=lambda coord: coord)
    print(nproc, out)


def test_two_gpus_large_shared_memory_kernels():
    """Per-device kernel attributes: two ranks as threads of one process, one GPU each, at a plane size whose
    kernels need > 48 KB of dynamic shared memory (the opt-in must have been made on BOTH devices)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    kw = dict(nx=512, ny=512, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, 2, nsteps=1, tol=1e-12, device_of=lambda coord: coord)
    print(out)
