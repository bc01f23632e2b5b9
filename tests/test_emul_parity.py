"""CPU-side checks of the product kernels' LOGIC: the same .cu sources compiled for the
host (tests/emul) must reproduce the oracle.  These run in the GPU-less container; the
real parity tests (tests/test_gpu_parity.py, -m gpu) run the sm_100a build on a B200."""
import shutil

import numpy as np
import pytest

import lesgo_b200
from helpers import (O, check_convec, check_derivatives, check_fft_raw, check_press, check_steps,
                     emul_library, make_dims)

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator")


def core_for(p):
    return lesgo_b200.Core(make_dims(p), lib=emul_library())


@pytest.mark.parametrize("nx,ny,Nz", [(16, 16, 4), (32, 16, 5), (48, 32, 3), (80, 48, 3)])
def test_emul_derivatives_and_raw_fft(nx, ny, Nz):
    p = O.Params(nx=nx, ny=ny, Nz=Nz, L_x=4.0, L_y=3.0)
    c = core_for(p)
    check_derivatives(c, p)
    check_fft_raw(c, p)


@pytest.mark.parametrize("bc", [(1, 1, False), (0, 0, False), (2, 2, True), (1, 0, True)])
def test_emul_convec(bc):
    p = O.Params(nx=16, ny=16, Nz=6, lbc_mom=bc[0], ubc_mom=bc[1], sgs=bc[2])
    check_convec(core_for(p), p)


def test_emul_misc():
    from helpers import check_misc
    p = O.Params(nx=16, ny=16, Nz=6, L_x=3.0)
    check_misc(core_for(p), p)


def test_emul_press():
    p = O.Params(nx=16, ny=16, Nz=8)
    check_press(core_for(p), p)


def test_emul_steps():
    p = O.Params(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5,
                 use_mean_p_force=True, mean_p_force_x=1.0)
    check_steps(core_for(p), p, nsteps=2, tol=1e-11)


@pytest.mark.parametrize("nproc", [2, 4])
def test_emul_multirank_steps(nproc):
    """z-slab ranks as threads over the in-process comm backend: ghost-plane sync, the
    slab->pencil transposes of the pressure solve and the k=0 chain."""
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(emul_library(), kw, nproc, nsteps=2)
    print(out)


@pytest.mark.parametrize("cfg", [
    dict(lbc_mom=1, ubc_mom=1, sgs=False, molec=True, nu_molec=1e-2, utop=1.0, ubot=-1.0),              # DNS Couette
    dict(lbc_mom=0, ubc_mom=0, sgs=False, molec=True, nu_molec=1e-2),                                    # stress-free DNS
    dict(lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False, use_mean_p_force=True, mean_p_force_x=1.0),   # LES channel
    dict(lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False),                                      # half channel, Cs = 0.03
    dict(lbc_mom=1, ubc_mom=1, sgs=True, sgs_model=1, molec=True, nu_molec=1e-3, ifilter=2),
])
def test_emul_full_step(cfg):
    """Rows (f)-1 on the device: wallstress, calc_Sij, constant-coefficient sgs_stag, divstress."""
    p = O.Params(nx=16, ny=16, Nz=8, **cfg)
    out = check_steps(core_for(p), p, nsteps=2, tol=1e-11, mode="full")
    print(out)


@pytest.mark.parametrize("nproc", [2, 4])
def test_emul_multirank_p2p_transposes(nproc):
    """Pressure transposes through peer memory (remote stores from the assembly and Thomas kernels)."""
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(emul_library(), kw, nproc, nsteps=3, p2p=True)
    print(out)


def test_emul_multirank_full_step():
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=8, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False,
              use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(emul_library(), kw, 2, nsteps=2, mode="full")
    print(out)


@pytest.mark.parametrize("cfg", [
    dict(nx=16, ny=16, Nz=8, lbc_mom=2, ubc_mom=0),               # half channel (LES_channel-like), equilibrium wall
    dict(nx=16, ny=32, Nz=6, lbc_mom=1, ubc_mom=1, molec=True),   # two DNS-type walls
    dict(nx=32, ny=16, Nz=6, lbc_mom=0, ubc_mom=2, ifilter=2),    # stress-free bottom, wall-modelled top, Gaussian filter
])
def test_emul_lasd_steps(cfg, monkeypatch):
    """Rows (f)-2: sgs_model 5 (lagrange_Sdep + interpolag_Sdep) over four steps with DYN_init = cs_count = 2,
    in plane chunks of 3 so the chunk seams are exercised."""
    from helpers import check_lasd_steps
    monkeypatch.setenv("LESGO_LASD_CHUNK", "3")
    p = O.Params(sgs=True, sgs_model=5, dt=4e-3, **cfg)
    out = check_lasd_steps(core_for(p), p, nsteps=4)
    print(out)


@pytest.mark.parametrize("nproc", [2, 4])
def test_emul_multirank_lasd(nproc):
    """The semi-Lagrangian transport reads ghost planes of F_* and u, v, w across slab seams."""
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=12, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, dt=4e-3)
    out = check_multirank_steps(emul_library(), kw, nproc, nsteps=4, mode="full", lasd=True)
    print(out)


@pytest.mark.parametrize("cfg,mode", [
    (dict(nx=32, ny=32, Nz=12, lbc_mom=1, ubc_mom=1), "core"),
    (dict(nx=32, ny=16, Nz=12, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False), "full"),
])
def test_emul_turbines(cfg, mode):
    """Rows (f)-3: actuator-disk forcing (turbines.f90:465-638) standalone and inside the step."""
    from helpers import check_turbines
    p = O.Params(**cfg)
    out = check_turbines(core_for(p), p, mode=mode, tol=1e-11)
    print(out)


def test_emul_turbines_overlapping_disks():
    from helpers import check_turbines
    p = O.Params(nx=32, ny=32, Nz=12, lbc_mom=1, ubc_mom=1)
    print(check_turbines(core_for(p), p, mode="core", tol=1e-11, overlap=True))


def test_emul_turbines_with_rotation():
    from helpers import check_turbines
    p = O.Params(nx=32, ny=32, Nz=12, lbc_mom=1, ubc_mom=1)
    print(check_turbines(core_for(p), p, mode="core", tol=1e-11, overlap=True, rotation=6.0))


@pytest.mark.parametrize("nproc", [2, 4])
def test_emul_multirank_turbines(nproc):
    """Disks spanning several z slabs: per-rank node lists, all-reduced disk velocities, force halos."""
    from helpers import check_multirank_steps
    kw = dict(nx=32, ny=16, Nz=12, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False)
    out = check_multirank_steps(emul_library(), kw, nproc, nsteps=2, mode="full", turbines=True, tavg=True,
                                rotation=6.0 if nproc == 4 else None)      # 4 ranks: the ADM with rotation
    print(out)


def test_emul_turbines_errors():
    p = O.Params(nx=16, ny=16, Nz=6)
    c = core_for(p)
    for n in ("u", "v", "w", "RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        c.upload(n, np.zeros(c.dims.shape))
    with pytest.raises(lesgo_b200.LibraryError, match="turbines_init"):
        c.step(dt=1e-3, turbines=True)
    bad = O.Turbine(xloc=1.0, yloc=1.0, height=0.5, dia=0.5, thk=0.1)
    bad.nodes = np.array([[1, 1, p.nz]], dtype=np.int32)      # k = nz is not a node a rank owns
    bad.ind = np.array([1.0])
    with pytest.raises(lesgo_b200.LibraryError, match="node outside"):
        c.turbines_init([bad])
    # use_rotation: the per-node arrays must follow an init call and match it
    import ctypes as C
    one = (C.c_void_p * 1)()
    with pytest.raises(lesgo_b200.LibraryError, match="turbines_init first"):
        c._ck(c.lib.turbines_rotation(c._ctx, 1, one, one, 7.0), "turbines_rotation")
    good = O.Turbine(xloc=1.0, yloc=1.0, height=0.5, dia=0.5, thk=0.1)
    good.nodes = np.array([[1, 1, 1], [2, 1, 1]], dtype=np.int32)
    good.ind = np.array([0.5, 0.5])
    c.turbines_init([good])
    with pytest.raises(lesgo_b200.LibraryError, match="must match"):
        c._ck(c.lib.turbines_rotation(c._ctx, 2, one, one, 7.0), "turbines_rotation")
    good.ind_t, good.e_theta = np.array([0.1, 0.2]), np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    with pytest.raises(lesgo_b200.LibraryError, match="tip_speed_ratio"):
        c.turbines_init([good], use_rotation=True, tip_speed_ratio=0.0)
    c.turbines_init([good], use_rotation=True, tip_speed_ratio=7.0)
    c.turbines_init([good])                  # handing the farm over again releases the old arrays and rotation


@pytest.mark.parametrize("cfg,turbines", [
    (dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, molec=True, nu_molec=1e-2), False),
    (dict(nx=32, ny=16, Nz=12, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False), True),
])
def test_emul_tavg(cfg, turbines):
    """Rows (f)-4: running time averages accumulated from the resident fields."""
    from helpers import check_tavg
    p = O.Params(**cfg)
    print(check_tavg(core_for(p), p, turbines=turbines, tol=1e-11))


def test_emul_checkpoint(tmp_path, monkeypatch):
    from helpers import check_checkpoint
    p = O.Params(nx=16, ny=16, Nz=6)
    check_checkpoint(lambda: core_for(p), p, tmp_path, monkeypatch)


@pytest.mark.parametrize("Nz,mode", [(2, "core"), (2, "full"), (3, "full")])
def test_emul_minimal_slab(Nz, mode):
    """Smallest slabs: bottom and top special planes adjacent (nz = 3 or 4)."""
    p = O.Params(nx=16, ny=16, Nz=Nz, lbc_mom=1, ubc_mom=1, utop=0.3, ubot=-0.3, sgs=(mode == "full"),
                 sgs_model=1, molec=True, nu_molec=1e-2)
    check_steps(core_for(p), p, nsteps=2, tol=1e-11, mode=mode)
    check_convec(core_for(p), p)


def test_emul_multirank_thin_slabs():
    """Four ranks with two owned planes each (nz = 3)."""
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, sgs=True, sgs_model=1,
              molec=True, nu_molec=1e-2)
    check_multirank_steps(emul_library(), kw, 4, nsteps=2, mode="full")


@pytest.mark.parametrize("env,grid,what", [
    ({"LESGO_PROD_CHUNK": "2"}, "32,16,7", "convec"),
    ({"LESGO_XW": "0"}, "16,16,6", "deriv,convec,steps"),
    ({"LESGO_XW": "1"}, "512,16,3", "deriv,convec"),
    ({"LESGO_XW": "3"}, "512,16,3", "deriv,convec,press"),
    ({"LESGO_XW": "2"}, "48,32,4", "deriv,convec,press,steps"),
    ({"LESGO_REUSE": "0"}, "16,16,6", "steps,full"),
    ({"LESGO_XW2": "7"}, "512,16,3", "deriv,convec,press,steps"),     # two-stage x inverse on every x-inverse launch
    ({"LESGO_XW2": "0"}, "512,16,3", "convec"),                       # ... and on none
])
def test_emul_variants(env, grid, what):
    """Kernel variants behind environment switches (read once per process -> subprocess)."""
    import os
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "variant_check.py"), "--emul", "--grid", grid, "--what", what],
                       env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "variant_check ok" in r.stdout, (env, r.stdout[-2000:], r.stderr[-2000:])


def test_emul_fftw_shim():
    """dfftw_plan_dft_r2c_2d / c2r_2d / execute / destroy_plan by reference, module fft's plans and
    turbine_indicator's out-of-place plans of another size (generic 2-3-5 transform)."""
    from helpers import check_fftw_shim
    p = O.Params(nx=16, ny=32, Nz=3)
    c = core_for(p)
    print(check_fftw_shim(emul_library(), c, p))


def test_emul_cfl_and_variable_dt():
    """get_max_cfl / get_cfl_dt on one slab, and a use_cfl_dt run (Euler start, tadv1 = 1 + dt/(2 dt_f))."""
    from helpers import check_cfl, check_variable_dt_steps
    p = O.Params(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5)
    check_cfl(core_for(p), p)
    print(check_variable_dt_steps(core_for(p), p, nsteps=4, tol=1e-11))
    p = O.Params(nx=16, ny=16, Nz=6, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False)
    print(check_variable_dt_steps(core_for(p), p, nsteps=3, tol=1e-11, mode="full"))


@pytest.mark.parametrize("p2p", [False, True])
def test_emul_multirank_ragged_ky_split(p2p):
    """ny not divisible by nproc (the reference has no such restriction): three ranks share 16 ky rows as 6 + 6 + 4."""
    from helpers import check_multirank_steps
    kw = dict(nx=16, ny=16, Nz=6, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5)
    print(check_multirank_steps(emul_library(), kw, 3, nsteps=2, p2p=p2p))
    kw = dict(nx=16, ny=16, Nz=9, lbc_mom=1, ubc_mom=0, sgs=True)     # three planes per rank: padded blocks outgrow a field
    print(check_multirank_steps(emul_library(), kw, 3, nsteps=2, p2p=p2p))
