"""Parity tests proper: the sm_100a CUDA path, called through the C ABI, against the
oracle on identical seeded inputs (north star: rel-L2 <= 1e-12 after one step, <= 1e-9
after ten, on u, v, w, p, RHS), plus size-independent properties at full size."""
import numpy as np
import pytest

import lesgo_b200
from helpers import (O, check_convec, check_derivatives, check_fft_raw, check_press, check_steps,
                     make_dims, random_field, rel)

pytestmark = pytest.mark.gpu


def core_for(p):
    return lesgo_b200.Core(make_dims(p, device=0))


@pytest.mark.parametrize("nx,ny,Nz", [(16, 16, 4), (32, 48, 5), (64, 64, 8), (128, 128, 8), (96, 80, 4),
                                      (160, 192, 3), (256, 256, 4), (320, 384, 2), (512, 512, 3),
                                      (1024, 512, 2), (384, 1024, 2)])
def test_derivatives_and_raw_fft(nx, ny, Nz):
    p = O.Params(nx=nx, ny=ny, Nz=Nz, L_x=4.0 * np.pi, L_y=2.0 * np.pi)
    c = core_for(p)
    check_derivatives(c, p)
    check_fft_raw(c, p, tol=2e-14)


@pytest.mark.parametrize("bc", [(1, 1, False), (0, 0, False), (2, 2, True), (1, 0, True)])
@pytest.mark.parametrize("nx,ny,Nz", [(32, 32, 6), (128, 64, 8)])
def test_convec(bc, nx, ny, Nz):
    p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=bc[0], ubc_mom=bc[1], sgs=bc[2])
    check_convec(core_for(p), p)


def test_convec_256():
    p = O.Params(nx=256, ny=256, Nz=6)
    check_convec(core_for(p), p)


@pytest.mark.parametrize("nx,ny,Nz", [(32, 32, 8), (128, 128, 64), (256, 128, 32)])
def test_press(nx, ny, Nz):
    p = O.Params(nx=nx, ny=ny, Nz=Nz)
    print(check_press(core_for(p), p, tol=1e-12))


def test_one_step_dns_couette_128x128x64():
    """BASELINE.json configs[1]: 128x128x64 DNS (no SGS) walls, 1e-12 after one step."""
    p = O.Params(nx=128, ny=128, Nz=64, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, L_x=4 * np.pi)
    out = check_steps(core_for(p), p, nsteps=1, tol=1e-12)
    print(out)


@pytest.mark.parametrize("nx,ny,Nz", [(96, 80, 6), (384, 320, 4), (160, 192, 5)])
def test_steps_on_mixed_radix_grids(nx, ny, Nz):
    """The 3 * 2**a and 5 * 2**a transform lengths (sizes.h) through the whole step -- convec's 3/2 grids are then
    144 x 120, 576 x 480, 240 x 288 -- core and full (Smagorinsky, wall model) modes, two steps."""
    p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, L_x=4.0, L_y=3.0, use_mean_p_force=True,
                 mean_p_force_x=1.0)
    print(check_steps(core_for(p), p, nsteps=2, tol=1e-11))
    p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False, L_x=4.0, L_y=3.0)
    print(check_steps(core_for(p), p, nsteps=2, tol=1e-11, mode="full"))


def test_ten_steps_1e9():
    p = O.Params(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0,
                 use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_steps(core_for(p), p, nsteps=10, tol=1e-9)
    print(out)


def test_stress_free_lid_step():
    p = O.Params(nx=64, ny=32, Nz=16, lbc_mom=1, ubc_mom=0)
    check_steps(core_for(p), p, nsteps=2, tol=1e-11)


def test_host_and_device_pointers_agree():
    """Host (numpy) arguments are staged; device (torch) arguments are used in place."""
    import torch
    p = O.Params(nx=64, ny=64, Nz=8)
    c = core_for(p)
    f = random_field(p, 5)
    hx, hy = c.empty(), c.empty()
    c.ddxy(f, hx, hy)
    tf = torch.from_numpy(f).cuda()
    tx, ty = torch.zeros_like(tf), torch.zeros_like(tf)
    c.ddxy(tf, tx, ty)
    c.synchronize()
    assert np.array_equal(tx.cpu().numpy()[:, :, :p.nx], hx[:, :, :p.nx])
    assert np.array_equal(ty.cpu().numpy()[:, :, :p.nx], hy[:, :, :p.nx])


def test_full_size_properties_512():
    """At BASELINE.json's full plane size the oracle is too slow for the whole grid, so use
    size-independent properties: r2c -> c2r round trip, filt_da idempotence, and the
    projected field of a step being divergence-free as the next step sees it."""
    p = O.Params(nx=512, ny=512, Nz=16, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0)
    c = core_for(p)
    f = random_field(p, 9)
    spec = c.fft_r2c(f.copy())
    back = c.fft_c2r(spec) / (p.nx * p.ny)
    assert rel(back[:, :, :p.nx], f[:, :, :p.nx]) < 1e-14
    g, gx, gy = f.copy(), c.empty(), c.empty()
    c.filt_da(g, gx, gy)
    h = g.copy()
    c.filt_da(h, gx, gy)
    assert rel(h[:, :, :p.nx], g[:, :, :p.nx]) < 1e-14
    check_steps(c, p, nsteps=1, tol=1e-12)           # all seven fields at the gate, full plane size
    # divergence of the projected field (rmsdiv.f90) after re-filtering, device-resident
    for n in ("u", "v", "w"):
        a = c.download(n)
        fa, fx, fy = a.copy(), c.empty(), c.empty()
        c.filt_da(fa, fx, fy)
        c.upload({"u": "dudx", "v": "dvdy", "w": "dwdz"}[n], fx if n == "u" else (fy if n == "v" else c.ddz_w(fa, c.empty())))
    d = c.rmsdiv()
    assert d < 1e-10, d


def test_errors_are_loud():
    with pytest.raises(lesgo_b200.LibraryError):
        lesgo_b200.Core(lesgo_b200.Dims(nx=100, ny=64, Nz=8, device=0))     # 100 not in the size list
    p = O.Params(nx=32, ny=32, Nz=4)
    c = core_for(p)
    with pytest.raises(lesgo_b200.LibraryError):
        c.ddx(np.zeros((3, 3, 3)), c.empty())


def test_full_step_dns_couette_128x128x64():
    """BASELINE.json configs[1] for real: DNS walls + molecular stress (wallstress, calc_Sij, sgs_stag
    with sgs = .false., divstress_uv/w on the device), 1e-12 after one step."""
    p = O.Params(nx=128, ny=128, Nz=64, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, L_x=4 * np.pi,
                 sgs=False, molec=True, nu_molec=1e-3)
    out = check_steps(core_for(p), p, nsteps=1, tol=1e-12, mode="full")
    print(out)


@pytest.mark.parametrize("cfg", [
    dict(lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False, use_mean_p_force=True, mean_p_force_x=1.0),
    dict(lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False),
    dict(lbc_mom=0, ubc_mom=0, sgs=False, molec=True, nu_molec=1e-2),
])
def test_full_step_les_channel(cfg):
    """LES channel with the equilibrium wall model and a constant-coefficient eddy viscosity
    (Smagorinsky, or the Cs_opt2 = 0.03 start-up phase of the dynamic models): 10 steps, 1e-9."""
    p = O.Params(nx=64, ny=64, Nz=32, **cfg)
    out = check_steps(core_for(p), p, nsteps=10, tol=1e-9, mode="full")
    print(out)


@pytest.mark.parametrize("cfg,nsteps,tol", [
    (dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, dt=2e-3), 10, 1e-9),     # LES half channel as LES_channel_Re1000
    (dict(nx=64, ny=48, Nz=16, lbc_mom=1, ubc_mom=1, molec=True, dt=2e-3), 4, 1e-11),
    (dict(nx=256, ny=256, Nz=8, lbc_mom=2, ubc_mom=2, ifilter=2, dt=1e-3), 2, 1e-11),   # BASELINE configs[2] plane size
])
def test_lasd_les_channel(cfg, nsteps, tol):
    """Rows (f)-2: Lagrangian scale-dependent dynamic model (lagrange_Sdep.f90 + interpolag_Sdep.f90) inside the
    full step, DYN_init = cs_count = 2; also checks the model's own state F_LM, F_MM, F_QN, F_NN, Cs_opt2."""
    from helpers import check_lasd_steps
    p = O.Params(sgs=True, sgs_model=5, **cfg)
    out = check_lasd_steps(core_for(p), p, nsteps=nsteps, tol=tol)
    print(out)


@pytest.mark.parametrize("cfg,mode", [
    (dict(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1), "core"),
    (dict(nx=128, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False), "full"),
])
def test_turbines(cfg, mode):
    """Rows (f)-3: actuator-disk forcing (turbines.f90:465-638): force fields, disk scalars, two steps."""
    from helpers import check_turbines
    p = O.Params(**cfg)
    out = check_turbines(core_for(p), p, mode=mode, tol=1e-12)
    print(out)


def test_turbines_overlapping_disks():
    from helpers import check_turbines
    p = O.Params(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1)
    print(check_turbines(core_for(p), p, mode="core", tol=1e-12, overlap=True))


def test_turbines_with_rotation():
    """use_rotation (turbines.f90:76, :607-615): the tangential force f_n e_theta ind_t / tip_speed_ratio, overlapping
    disks included (the last disk's assignment wins, rotation term and all)."""
    from helpers import check_turbines
    p = O.Params(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1)
    print(check_turbines(core_for(p), p, mode="core", tol=1e-12, overlap=True, rotation=6.0))


@pytest.mark.parametrize("cfg,turbines", [
    (dict(nx=64, ny=64, Nz=16, lbc_mom=1, ubc_mom=1, molec=True, nu_molec=1e-2), False),
    (dict(nx=64, ny=32, Nz=24, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False), True),
])
def test_tavg(cfg, turbines):
    """Rows (f)-4: tavg%compute (time_average.f90:176-320) from the resident fields, all 26 accumulators."""
    from helpers import check_tavg
    p = O.Params(**cfg)
    print(check_tavg(core_for(p), p, turbines=turbines, tol=1e-12))


def test_checkpoint(tmp_path, monkeypatch):
    """Rows (f)-4: restart file of the resident state (io.f90:1204-1211 / initial.f90:226-239)."""
    from helpers import check_checkpoint
    p = O.Params(nx=64, ny=32, Nz=12)
    check_checkpoint(lambda: core_for(p), p, tmp_path, monkeypatch)


def test_misc_entry_points():
    from helpers import check_misc
    p = O.Params(nx=64, ny=32, Nz=12, L_x=3.0)
    c = core_for(p)
    check_misc(c, p)
    # profiling facility + launch counter
    f = random_field(p, 3)
    n0 = c.launch_count
    c.profile(True)
    c.ddx(f, c.empty())
    rep = c.profile(False, report=True)
    assert c.launch_count - n0 == 3 and set(rep) == {"xfwd", "ypass_deriv", "xinv"}


# ---- round 2: the benchmarked configurations themselves, variable dt, the FFTW boundary, many slabs on one GPU ----
def test_full_size_one_step_512x512x256():
    """BASELINE.json's headline grid, whole: one core step of 512 x 512 x 256 against the oracle, all seven
    fields at the north star's gate (rel-L2 <= 1e-12).  The oracle takes about two minutes of host time."""
    p = O.Params(nx=512, ny=512, Nz=256, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0)
    out = check_steps(core_for(p), p, nsteps=1, tol=1e-12)
    print(out)


def test_cfl_single_slab():
    from helpers import check_cfl
    p = O.Params(nx=64, ny=48, Nz=16, L_x=4.0, L_y=3.0)
    print(check_cfl(core_for(p), p))


@pytest.mark.parametrize("cfg,mode", [
    (dict(nx=64, ny=64, Nz=32, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0), "core"),
    # LES_channel_Re1000 as shipped: half channel, wall model below, stress-free lid, use_cfl_dt, cfl = 0.0625
    (dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False,
          use_mean_p_force=True, mean_p_force_x=1.0), "full"),
])
def test_variable_dt_ten_steps(cfg, mode):
    """use_cfl_dt (lesgo.conf:117 of the shipped LES_channel_Re1000): dt = get_cfl_dt() every step, tadv1 = 1 +
    dt/(2 dt_f), Euler start -- main.f90:135-144, initialize.f90:192-199, cfl_util.f90:72-113."""
    from helpers import check_variable_dt_steps
    p = O.Params(**cfg)
    out = check_variable_dt_steps(core_for(p), p, nsteps=10, tol=1e-9, mode=mode)
    print(out)


def test_variable_dt_one_step_gate():
    from helpers import check_variable_dt_steps
    p = O.Params(nx=128, ny=128, Nz=64, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, L_x=4 * np.pi)
    print(check_variable_dt_steps(core_for(p), p, nsteps=1, tol=1e-12))


def test_fftw_shim_symbols():
    """dfftw_plan_dft_r2c_2d_ ... dfftw_destroy_plan_ called by reference as gfortran does: module fft's in-place
    plans (fft.f90:114-121) and turbine_indicator.f90:130-151's own 2048 x 2048 out-of-place plans."""
    from helpers import check_fftw_shim
    p = O.Params(nx=64, ny=48, Nz=3)
    c = core_for(p)
    print(check_fftw_shim(lesgo_b200.load_library(), c, p, big_generic=(2048, 2048)))


@pytest.mark.parametrize("nproc,p2p", [(2, False), (4, True), (8, False), (8, True), (3, True)])
def test_many_slabs_on_one_gpu(nproc, p2p):
    """The multi-slab path on ONE device (single-device transport, comm.cu): nproc z-slab ranks as threads of this
    process, all on cuda:0 -- ghost-plane halos (mpi_defs.f90:245-262), the slab <-> pencil transposes that replace
    tridag_array.f90:85-157, the k = 0 chain (press_stag_array.f90:221-244), cfl all-reduce -- against the
    single-slab oracle.  nproc = 3 splits the 64 ky rows raggedly (22 + 22 + 20)."""
    from helpers import check_multirank_steps
    kw = dict(nx=64, ny=64, Nz=24, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, use_mean_p_force=True, mean_p_force_x=1.0)
    out = check_multirank_steps(lesgo_b200.load_library(), kw, nproc, nsteps=2, tol=1e-12, p2p=p2p, local=True,
                                device_of=lambda coord: 0)
    print(nproc, out)


def test_many_slabs_on_one_gpu_full_models():
    """Same, with everything that communicates: wall model + Lagrangian scale-dependent model (F_* halos), actuator
    disks (device all-reduce of the disk velocities) and the running averages (interpolation halos)."""
    from helpers import check_multirank_steps
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, dt=2e-3)
    print(check_multirank_steps(lesgo_b200.load_library(), kw, 4, nsteps=4, tol=1e-11, mode="full", lasd=True, local=True,
                                device_of=lambda coord: 0))
    kw = dict(nx=64, ny=64, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False)
    print(check_multirank_steps(lesgo_b200.load_library(), kw, 4, nsteps=2, tol=1e-11, mode="full", turbines=True, tavg=True,
                                local=True, p2p=True, device_of=lambda coord: 0, rotation=6.0))


def test_eight_slabs_512x512_planes_on_one_gpu():
    """BASELINE configs[3]'s split (eight slabs) at the full 512 x 512 plane size, 32 levels: the >48 KB dynamic
    shared-memory kernels on every rank's context, two steps at the 1e-12 gate."""
    from helpers import check_multirank_steps
    kw = dict(nx=512, ny=512, Nz=32, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0)
    print(check_multirank_steps(lesgo_b200.load_library(), kw, 8, nsteps=1, tol=1e-12, p2p=True, local=True,
                                device_of=lambda coord: 0))


def test_adm_1024x512_planes():
    """BASELINE configs[4] (turbines_ADM, 1024 x 512, 3/2 grid 1536 x 768): convec on that plane size and two full
    LES steps with actuator disks."""
    from helpers import check_turbines
    p = O.Params(nx=1024, ny=512, Nz=8, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False, L_x=4 * np.pi)
    check_convec(core_for(p), p)
    print(check_turbines(core_for(p), p, mode="full", tol=1e-12))


def test_lasd_256x256x32():
    """BASELINE configs[2] (256 x 256 LES channel, Lagrangian scale-dependent model), 32 levels, four steps with two
    model updates."""
    from helpers import check_lasd_steps
    p = O.Params(nx=256, ny=256, Nz=32, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, dt=1e-3)
    print(check_lasd_steps(core_for(p), p, nsteps=4, tol=1e-11))
