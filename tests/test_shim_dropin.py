"""The drop-in boundary, executed: the reference's own sources (main.f90's time loop, wallstress, sgs_stag, divstress ...,
interpreted by oracle/f90exec.py where they lie under /root/reference) with the five sources of the hot path replaced by the
ISO_C_BINDING shims of fortran/ -- derivatives.f90, convec.f90, press_stag_array.f90, tridag_array.f90, fft.f90 -- whose
bind(C) calls go through ctypes into liblesgo_cuda's kernel-logic build (tests/shim_driver.py).  The results must equal what
the ALL-reference run produced (tests/golden/ref_*.npz).  Needs the reference tree, so it runs in the build container only
(no Fortran compiler exists on any box of this pool: this is the only way the shims get executed)."""
import ast
import os
import shutil

import numpy as np
import pytest

from helpers import O, make_dims, rel, emul_library
import lesgo_b200
from oracle import refrun

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.skipif(not refrun.available() or shutil.which("g++") is None,
                                reason="needs the reference sources under /root/reference and g++ for the emulator")
FIELDS = ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz")
refrun.MPI_TIMEOUT = min(refrun.MPI_TIMEOUT, 120.0)     # a rank that fails must not keep its peers waiting for ten minutes


def shimmed(p, **kw):
    from shim_driver import ShimmedReference
    return ShimmedReference(p, lesgo_b200.Core(make_dims(p), lib=emul_library()), **kw)


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return d, ast.literal_eval(str(d["meta"]))


def test_routines_through_the_fortran_shims():
    """ddx, ddy, ddxy, filt_da, ddz_uv, ddz_w called as main.f90 / divstress call them (module arrays of sim_param as
    actual arguments, assumed-shape dummies in the shim), convec for four wall-condition combinations and
    press_stag_array (the shim hands dpdx, dpdy, dpdz over shifted down by one plane): vs the reference's own routines."""
    from helpers import random_field
    d, meta = load("ref_routines_16x16x6")
    kw = dict(nx=16, ny=16, Nz=6, L_x=4.0, L_y=3.0)
    p = O.Params(**kw)
    R = shimmed(p)
    I, D = R.I, "derivatives"
    S = lambda n: I.get("sim_param", n)
    R.put("u", d["f"])
    I.call("ddx", S("u"), S("dudx"), 0, module=D)
    I.call("ddy", S("u"), S("dudy"), 0, module=D)
    I.call("ddxy", S("u"), S("dvdx"), S("dvdy"), 0, module=D)
    I.call("ddz_uv", S("u"), S("dudz"), 0, module=D)
    I.call("ddz_w", S("u"), S("dwdz"), 0, module=D)
    got = {"ddx": R.get("dudx"), "ddy": R.get("dudy"), "ddxy_x": R.get("dvdx"), "ddxy_y": R.get("dvdy"),
           "ddz_uv": R.get("dudz"), "ddz_w": R.get("dwdz")}
    I.call("filt_da", S("u"), S("dudx"), S("dudy"), 0, module=D)
    got.update(filt_da_f=R.get("u"), filt_da_x=R.get("dudx"), filt_da_y=R.get("dudy"))
    nz, nx = p.nz, p.nx
    for n, g in got.items():
        lo = 2 if n == "ddz_uv" else 1            # ddz_uv leaves plane 1 of the bottom rank to the wall model
        hi = nz if n != "ddz_w" else nz - 1
        assert rel(g[lo:hi, :, :nx], d[n][lo:hi, :, :nx]) <= 1e-14, n
    for tag, bc in (("11d", (1, 1, False)), ("00d", (0, 0, False)), ("22l", (2, 2, True)), ("10l", (1, 0, True))):
        pc = O.Params(lbc_mom=bc[0], ubc_mom=bc[1], sgs=bc[2], **kw)
        Rc = shimmed(pc, files=[x for x in refrun.FILES if x not in ("sgs_stag_util.f90", "wallstress.f90", "divstress_uv.f90", "divstress_w.f90")])
        for i, n in enumerate(("u", "v", "w", "dudy", "dudz", "dvdx", "dvdz", "dwdx", "dwdy")):
            Rc.put(n, random_field(pc, 20 + i))
        Rc.call("convec")
        assert Rc.calls.get("lesgo_gpu_convec") == 1
        for n in ("RHSx", "RHSy", "RHSz"):
            hi = pc.nz + 1 if n == "RHSz" and False else pc.nz
            assert rel(Rc.get(n)[1:hi, :, :nx], d[f"convec_{tag}_{n}"][1:hi, :, :nx]) <= 1e-13, (tag, n)
    Rp = shimmed(O.Params(**kw))
    for n in ("u", "v", "w", "divtz"):
        Rp.put(n, d["press_" + n])
    Rp.call("press_stag_array")
    assert Rp.calls.get("lesgo_gpu_press_stag_array") == 1
    for n in ("p", "dpdx", "dpdy", "dpdz"):
        hi = nz + 1 if n == "p" else nz
        lo = 0 if n == "p" else 1
        assert rel(Rp.get(n)[lo:hi, :, :nx], d["press_" + n][lo:hi, :, :nx]) <= 1e-12, n


@pytest.mark.parametrize("name", ["ref_core_couette_32x32x8", "ref_full_couette_16x16x8", "ref_full_les_channel_16x32x8",
                                  "ref_core_free_slip_les_16x16x6"])
def test_reference_time_loop_over_the_fortran_shims(name):
    """main.f90:155-344 from the reference text -- with wallstress, calc_Sij, sgs_stag, divstress_uv / divstress_w still the
    reference's Fortran, calling ddx / ddy / ddxy / ddz_* of the shim -- over the shimmed filt_da, convec and
    press_stag_array: the fields after one step and after the fixture's last step vs the all-reference run."""
    d, meta = load(name)
    p = O.Params(**meta["params"])
    R = shimmed(p)
    for n in ("u", "v", "w"):
        R.put(n, d[n + "0"])
    last = min(max(meta["record"]), 4)            # the interpreter runs ~3 s per step: four steps are enough here
    for it in range(1, last + 1):
        R.step(it, mode=meta["mode"])
        if it == 1:
            for n in FIELDS:
                hi = p.nz + 1 if n in ("w", "RHSz", "p") else p.nz
                assert rel(R.get(n)[1:hi, :, :p.nx], d[f"{n}_1"][1:hi, :, :p.nx]) <= 1e-13, (n, it)
    assert R.calls["lesgo_gpu_filt_da"] == 3 * last and R.calls["lesgo_gpu_convec"] == last
    assert R.calls["lesgo_gpu_press_stag_array"] == last
    if last in meta["record"]:
        for n in FIELDS:
            hi = p.nz + 1 if n in ("w", "RHSz", "p") else p.nz
            assert rel(R.get(n)[1:hi, :, :p.nx], d[f"{n}_{last}"][1:hi, :, :p.nx]) <= 1e-12, (n, last)


def test_tridag_array_shim():
    """The explicit-shape shim of tridag_array (kept for callers other than press_stag_array): a diagonally dominant
    system per (kx, ky) mode vs numpy."""
    p = O.Params(nx=16, ny=16, Nz=6)
    R = shimmed(p)
    I = R.I
    rng = np.random.default_rng(5)
    n = p.nz + 1
    a = rng.uniform(-1, 1, (n, p.ny, p.lh)); c = rng.uniform(-1, 1, (n, p.ny, p.lh)); b = 4.0 + rng.uniform(0, 1, (n, p.ny, p.lh))
    r = rng.uniform(-1, 1, (n, p.ny, p.ld))
    from oracle import f90exec as F
    fa = lambda x: F.FArray(np.asfortranarray(x.transpose(2, 1, 0).copy()), (1, 1, 1))
    A, B, Cc, Rr, U = fa(a), fa(b), fa(c), fa(r), fa(np.zeros_like(r))
    I.call("tridag_array", A, B, Cc, Rr, U)
    assert R.calls.get("lesgo_gpu_tridag_array") == 1
    u = U.a.transpose(2, 1, 0)
    # solve mode (jx, jy) for the real and the imaginary part: rows 1..n of a x(j-1) + b x(j) + c x(j+1) = r(j)
    for jy in (0, 3, p.ny - 1):
        for jx in (1, 4, p.lh - 2):
            M = np.diag(b[:, jy, jx]) + np.diag(a[1:, jy, jx], -1) + np.diag(c[:-1, jy, jx], 1)
            for part in (0, 1):
                x = np.linalg.solve(M, r[:, jy, 2 * jx + part])
                assert np.allclose(u[:, jy, 2 * jx + part], x, rtol=1e-11, atol=1e-13), (jx, jy, part)


def _step_params(R, p, dt, t1, t2, first):
    from oracle import f90exec as F
    sp = F.FStruct(R.I.modules["lesgo_gpu_resident_mod"].types["lesgo_gpu_step_params"])
    vals = dict(dt=dt, tadv1=t1, tadv2=t2, mean_p_force_x=p.mean_p_force_x if p.use_mean_p_force else 0.0,
                mean_p_force_y=p.mean_p_force_y if p.use_mean_p_force else 0.0, ubot=p.ubot, utop=p.utop, nu_molec_nd=p.nu,
                first_step=int(first), mode=1, sgs_model=p.sgs_model, ifilter=p.ifilter, co=p.Co, wall_damp_exp=p.wall_damp_exp,
                vonk=p.vonk, zo=p.zo, lasd_cs_init=0, lasd_update=0, lasd_init_f=0, lagran_dt=0.0, turbines=0, turbines_eps=0.0)
    for k, v in vals.items():
        setattr(sp, k, v)
    return sp


@pytest.mark.parametrize("use_cfl_dt", [True, False])
def test_resident_step_with_the_fortran_lasd_switches(use_cfl_dt):
    """fortran/lesgo_gpu_resident_mod.f90: gpu_lasd_switches -- the Fortran a maintainer calls before lesgo_gpu_step with
    sgs_model 5 -- interpreted statement by statement with the reference's own counters (jt, jt_total, DYN_init, cs_count,
    dt, use_cfl_dt of module param), filling the bind(C) lesgo_gpu_step_params the library then steps with: against the
    reference-source fixtures of the Lagrangian model with a varying (accumulated lagran_dt) and a fixed
    time step."""
    from oracle import f90exec as F
    from helpers import LasdClock
    name = "ref_full_lasd_cfl_dt_16x16x6" if use_cfl_dt else "ref_full_lasd_16x16x6"
    d, meta = load(name)
    p = O.Params(**meta["params"])
    R = shimmed(p, files=refrun.LASD_FILES, resident=True, dyn_init=meta["dyn_init"], cs_count=meta["cs_count"])
    I, core = R.I, R.core
    for n in ("u", "v", "w"):
        core.upload(n, d[n + "0"])
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz", "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
        core.upload(n, np.zeros(core.dims.shape))
    I.set("param", "use_cfl_dt", bool(use_cfl_dt))
    flag = F.FArray(np.zeros(1, dtype=bool), (1,), "logical")
    clock = LasdClock(meta["cs_count"], meta["dyn_init"], use_cfl_dt)
    last = max(meta["record"])
    for it in range(1, last + 1):
        dt, t1, t2 = (float(x) for x in d["dts"][it - 1]) if use_cfl_dt else (p.dt, p.tadv1, p.tadv2)
        I.set("param", "jt", it); I.set("param", "jt_total", it); I.set("param", "dt", dt)
        sp = _step_params(R, p, dt, t1, t2, first=(it == 1))
        I.call("gpu_lasd_switches", sp, F.ElemRef(flag, (0,)), module="lesgo_gpu_resident_mod")
        want = clock.switches(it, dt)
        got = dict(lasd_cs_init=bool(sp.lasd_cs_init), lasd_update=bool(sp.lasd_update), lasd_init_F=bool(sp.lasd_init_f),
                   lagran_dt=float(sp.lagran_dt))
        assert got == want, (it, got, want)
        rc = I.externals["lesgo_gpu_step"](None, [core._ctx_ptr(), sp])
        assert rc == 0, core.lib.error(core._ctx)
        if it in meta["record"]:
            for n in FIELDS + ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
                top = n in ("w", "RHSz", "p", "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2")
                hi = p.nz + 1 if top else p.nz
                e = rel(core.download(n)[1:hi, :, :p.nx], d[f"{n}_{it}"][1:hi, :, :p.nx])
                assert e <= (1e-10 if n == "Cs_opt2" else 1e-11), (n, it, e)
    assert bool(flag.a[0])


@pytest.mark.parametrize("fixture", ["ref_turbines_32x32x8", "ref_turbines_rot_32x32x8"])
def test_fortran_wind_farm_hand_over(fixture):
    """fortran/lesgo_gpu_resident_mod.f90: gpu_turbines_set, interpreted: it walks the reference's own wind_farm (stat_defs'
    derived types, as turbines_nodes leaves it), transposes %nodes and %e_theta into C order, takes c_loc of the members and
    calls lesgo_gpu_turbines_init (+ lesgo_gpu_turbines_rotation with use_rotation).  The library stepped afterwards must
    reproduce the reference-source actuator-disk fixtures."""
    from test_reference_pin import farm_from_fixture, rotation_kw
    from helpers import step_kwargs_pre_dyn
    d, meta = load(fixture)
    p = O.Params(**meta["params"])
    R = shimmed(p, files=refrun.TURBINE_FILES, resident=True, turbines=True)
    I, core = R.I, R.core
    farm = farm_from_fixture(d, meta)
    rk = rotation_kw(meta)
    R.farm_set(farm, meta["eps"], adm_correction=meta["adm_correction"], **rk)
    I.call("gpu_turbines_set", I.get("stat_defs", "wind_farm"), len(farm), bool(meta["adm_correction"]), rk["use_rotation"],
           rk["tip_speed_ratio"], module="lesgo_gpu_resident_mod")
    assert R.calls.get("lesgo_gpu_turbines_init") == 1
    assert R.calls.get("lesgo_gpu_turbines_rotation", 0) == (1 if rk["use_rotation"] else 0)
    for n in ("u", "v", "w"):
        core.upload(n, d[n + "0"])
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(core.dims.shape))
    n = meta["nsteps"]
    for it in range(1, n + 1):
        core.step(**step_kwargs_pre_dyn(p, it - 1, "core"), turbines=True, turbines_eps=meta["eps"])
    for name in ("fxa", "fya", "fza"):
        assert rel(core.download(name)[1:p.nz, :, :p.nx], d[f"{name}_{n}"][1:p.nz, :, :p.nx]) <= 1e-12, name
    for name in FIELDS:
        hi = p.nz + 1 if name in ("w", "RHSz", "p") else p.nz
        assert rel(core.download(name)[1:hi, :, :p.nx], d[f"{name}_{n}"][1:hi, :, :p.nx]) <= 1e-12, name


def test_the_shims_own_start_up():
    """Nothing pre-made: init_fft of fortran/fft.f90 calls gpu_require, which builds lesgo_gpu_dims with a structure
    constructor from module param and calls lesgo_gpu_create; gpu_pin_sim_param page-locks the twenty sim_param arrays;
    gpu_check is the Fortran one.  One core step afterwards equals the all-reference run."""
    import ctypes as C
    from shim_driver import ShimmedReference
    d, meta = load("ref_core_couette_32x32x8")
    p = O.Params(**meta["params"])
    lib = emul_library()
    R = ShimmedReference(p, core=None, lib=lib)
    try:
        assert R.ctx is not None and R.ctx.addr != 0
        assert R.calls["lesgo_gpu_create"] == 1 and R.calls["lesgo_gpu_host_register"] == 20, R.calls
        for n in ("u", "v", "w"):
            R.put(n, d[n + "0"])
        R.step(1, mode="core")
        assert R.calls["lesgo_gpu_create"] == 1          # c_associated(gpu_ctx): created once
        for n in FIELDS:
            hi = p.nz + 1 if n in ("w", "RHSz", "p") else p.nz
            assert rel(R.get(n)[1:hi, :, :p.nx], d[f"{n}_1"][1:hi, :, :p.nx]) <= 1e-13, n
    finally:
        if R.ctx is not None and R.ctx.addr:
            lib.destroy(C.c_void_p(R.ctx.addr))


def test_four_mpi_ranks_over_the_shims():
    """`mpirun -np 4` in miniature: four interpreted ranks of the reference (its own mpi_sync_real_array halos over the
    interpreter's mailboxes) with the shims' own multi-rank start-up -- gpu_require makes the transport id on coord 0,
    MPI_Bcasts it, initialises the library's communicator, exports / all-gathers / imports the peer-memory blobs with the
    all-reduced success flags -- and the library's slab <-> pencil pressure solve in place of the pipelined tridag_array:
    two full steps, gathered, vs the reference's own four-rank run."""
    import ctypes as C
    import threading
    from shim_driver import ShimmedReference
    d = np.load(os.path.join(GOLD, "ref_mpi4_full_16x16x8.npz"))
    meta = ast.literal_eval(str(d["meta"]))
    kw, nproc, nsteps = meta["kw"], meta["nproc"], meta["nsteps"]
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=meta["seed"], amp=meta["amp"], L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)
    lib = emul_library()
    boxes = {"lock": threading.Lock()}
    res, err, refs = [None] * nproc, [None] * nproc, [None] * nproc

    def work(r):
        try:
            p = O.Params(nproc=nproc, coord=r, **kw)
            R = refs[r] = ShimmedReference(p, core=None, lib=lib, boxes=boxes)
            for n, g in (("u", ug), ("v", vg), ("w", wg)):
                R.put(n, O.scatter_slab(g, p))
            for it in range(1, nsteps + 1):
                R.step(it, mode=meta["mode"])
            res[r] = {n: R.get(n) for n in FIELDS}
        except BaseException as e:  # noqa
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nproc)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    try:
        real = [e for e in err if e is not None and type(e).__name__ != "Empty"]
        if real:
            raise real[0]
        assert not any(err), err
        for r in range(nproc):
            c = refs[r].calls
            assert c["lesgo_gpu_comm_init"] == 1 and c["lesgo_gpu_comm_p2p_export"] == 1 and c["lesgo_gpu_comm_p2p_import"] == 1, c
            assert c.get("lesgo_gpu_comm_unique_id", 0) == (1 if r == 0 else 0), c
            assert c["lesgo_gpu_press_stag_array"] == nsteps
        ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
        for n in FIELDS:
            top = n in ("w", "RHSz", "p")
            g = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=top)
            hi = pg.nz_tot if top else pg.nz_tot - 1
            assert rel(g[1:hi + 1, :, :pg.nx], d[n][1:hi + 1, :, :pg.nx]) <= 1e-12, n
    finally:
        for R in refs:
            if R is not None and R.ctx is not None and R.ctx.addr:
                lib.destroy(C.c_void_p(R.ctx.addr))


def test_padd_unpadd_shims_against_the_reference_routines():
    """padd / unpadd of fortran/fft.f90 (one plane, explicit-shape dummies) vs the reference's own fft.f90:43-99, both
    interpreted in this process on the same random plane."""
    from oracle import f90exec as F
    from helpers import random_field
    p = O.Params(nx=16, ny=16, Nz=4, L_x=4.0, L_y=3.0)
    ref = refrun.Reference(p)
    shim = shimmed(p)
    rng = np.random.default_rng(11)
    small = rng.standard_normal((p.ny, p.ld))
    small[:, p.ld - 2:] = 0.0
    big = rng.standard_normal((p.ny2, p.ld_big))
    fa = lambda x: F.FArray(np.asfortranarray(x.T.copy()), (1, 1))
    out = {}
    for tag, R in (("ref", ref), ("shim", shim)):
        ub, u = fa(np.zeros((p.ny2, p.ld_big))), fa(small)
        R.I.call("padd", ub, u, module="fft")
        cc, cb = fa(np.zeros((p.ny, p.ld))), fa(big)
        R.I.call("unpadd", cc, cb, module="fft")
        out[tag] = (ub.a.T.copy(), cc.a.T.copy())
    assert shim.calls.get("lesgo_gpu_padd") == 1 and shim.calls.get("lesgo_gpu_unpadd") == 1
    assert np.count_nonzero(out["ref"][0]) > 50 and np.array_equal(out["shim"][0], out["ref"][0])
    assert np.array_equal(out["shim"][1][:, :p.ld - 2], out["ref"][1][:, :p.ld - 2])


def test_gpu_test_filter_wrapper_against_the_reference_test_filter():
    """gpu_test_filter(f, G_test) / (f, G_test_test) of fortran/lesgo_gpu_mod.f90 -- the one-line bodies of test_filter and
    test_test_filter -- vs the reference's own routines (test_filtermodule.f90:126-168) on the same plane, with the
    reference's own kernels, for the three filter types."""
    from oracle import f90exec as F
    from helpers import random_field
    for ifilter in (1, 2, 3):
        p = O.Params(nx=16, ny=32, Nz=4, L_x=4.0, L_y=3.0, sgs=True, sgs_model=5, ifilter=ifilter)
        R = shimmed(p, files=refrun.LASD_FILES)
        I = R.I
        f = random_field(p, 7)[2]
        for kernel, routine in (("g_test", "test_filter"), ("g_test_test", "test_test_filter")):
            a = F.FArray(np.asfortranarray(f.T.copy()), (1, 1))
            I.call(routine, a, module="test_filtermodule")                    # the reference's body (FFTW -> pocketfft)
            b = F.FArray(np.asfortranarray(f.T.copy()), (1, 1))
            I.call("gpu_test_filter", b, I.get("test_filtermodule", kernel), module="lesgo_gpu_mod")
            assert rel(b.a.T[:, :p.nx], a.a.T[:, :p.nx]) <= 1e-14, (ifilter, routine)
        assert R.calls.get("lesgo_gpu_test_filter") == 2


def test_gpu_sync_real_array_wrapper_on_two_ranks():
    """gpu_sync_real_array(var, isync) of fortran/lesgo_gpu_mod.f90 over the library's transport vs the reference's
    mpi_sync_real_array (mpi_defs.f90:167-264) over the interpreter's MPI, on two ranks, for DOWN, UP and DOWNUP."""
    import ctypes as C
    import threading
    from oracle import f90exec as F
    from shim_driver import ShimmedReference
    kw = dict(nx=16, ny=16, Nz=8, L_x=4.0, L_y=3.0)
    nproc = 2
    lib = emul_library()
    boxes = {"lock": threading.Lock()}
    err, refs, out = [None] * nproc, [None] * nproc, [None] * nproc

    def work(r):
        try:
            p = O.Params(nproc=nproc, coord=r, **kw)
            R = refs[r] = ShimmedReference(p, core=None, lib=lib, boxes=boxes)
            I = R.I
            got = {}
            for isync in (1, 2, 3):
                base = np.arange((p.nz + 1) * p.ny * p.ld, dtype=np.float64).reshape(p.nz + 1, p.ny, p.ld) + 1000.0 * r
                a = F.FArray(np.asfortranarray(base.transpose(2, 1, 0).copy()), (1, 1, 0))
                I.call("mpi_sync_real_array", a, 0, isync, module="mpi_defs")
                b = F.FArray(np.asfortranarray(base.transpose(2, 1, 0).copy()), (1, 1, 0))
                I.call("gpu_sync_real_array", b, isync, module="lesgo_gpu_mod")
                got[isync] = (a.a.copy(), b.a.copy())
            out[r] = got
        except BaseException as e:  # noqa
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nproc)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    try:
        real = [e for e in err if e is not None and type(e).__name__ != "Empty"]
        if real:
            raise real[0]
        assert not any(err), err
        for r in range(nproc):
            for isync, (a, b) in out[r].items():
                assert np.array_equal(a, b), (r, isync)
            assert refs[r].calls.get("lesgo_gpu_sync_real_array") == 3
        # and the halos really moved: plane nz of rank 0 is plane 1 of rank 1 after SYNC_DOWN
        assert np.array_equal(out[0][1][1][:, :, -1], out[1][1][1][:, :, 1])
    finally:
        for R in refs:
            if R is not None and R.ctx is not None and R.ctx.addr:
                lib.destroy(C.c_void_p(R.ctx.addr))


def test_gpu_checkpoint_wrapper(tmp_path):
    """gpu_checkpoint(fname, writing) of fortran/lesgo_gpu_resident_mod.f90: the file name goes over as
    trim(fname) // c_null_char; the restart record it writes is the oracle writer's byte for byte (io.f90:1204-1211), and
    reading it back through the wrapper restores the resident fields."""
    from helpers import random_field
    p = O.Params(nx=16, ny=16, Nz=6)
    R = shimmed(p, resident=True)
    core = R.core
    s = O.State(p)
    O.lasd_alloc(s)
    for i, n in enumerate(O.CHECKPOINT_FIELDS):
        getattr(s, n)[...] = random_field(p, 300 + i)
        core.upload(n, getattr(s, n))
    f1, f2 = str(tmp_path / "vel.out.c0"), str(tmp_path / "vel.oracle.c0")
    R.I.call("gpu_checkpoint", f1 + "   ", True, module="lesgo_gpu_resident_mod")       # trailing blanks as a character(*) has
    O.checkpoint_write(s, p, f2)
    assert open(f1, "rb").read() == open(f2, "rb").read()
    for n in O.CHECKPOINT_FIELDS:
        core.upload(n, np.zeros(core.dims.shape))
    R.I.call("gpu_checkpoint", f2, False, module="lesgo_gpu_resident_mod")
    for n in O.CHECKPOINT_FIELDS:
        assert np.array_equal(core.download(n)[1:], getattr(s, n)[1:]), n
    assert R.calls["lesgo_gpu_checkpoint_write"] == 1 and R.calls["lesgo_gpu_checkpoint_read"] == 1


def test_gpu_check_carries_the_library_message_into_the_reference_error_routine():
    """gpu_check of fortran/lesgo_gpu_mod.f90 on a failing call: it fetches lesgo_gpu_last_error, turns the C string into a
    Fortran one with c_f_pointer and hands it to the reference's own fatal-error routine (messages.f90)."""
    import ctypes as C
    from oracle import f90exec as F
    from shim_driver import ShimmedReference
    p = O.Params(nx=16, ny=16, Nz=4)
    lib = emul_library()
    R = ShimmedReference(p, core=None, lib=lib)
    try:
        plane = np.zeros((p.ny, p.ld))
        rc = lib.test_filter(C.c_void_p(R.ctx.addr), plane.ctypes.data, plane.ctypes.data, 10 ** 6)
        assert rc != 0
        with pytest.raises(F.FStop, match="messages.f90"):                    # the reference's error(): writes, then `stop 1`
            R.I.call("gpu_check", rc, "test_filter", module="lesgo_gpu_mod")
        printed = "\n".join(R.I.written[-5:])
        assert "lesgo_gpu.test_filter" in printed and "at most nz" in printed, printed
        R.I.call("gpu_check", 0, "anything", module="lesgo_gpu_mod")          # rc = 0 returns quietly
    finally:
        lib.destroy(C.c_void_p(R.ctx.addr))


def test_reference_lagrangian_model_over_the_fortran_shims():
    """The same with sgs_model 5: lagrange_Sdep, interpolag_Sdep, test_filter ... stay the reference's Fortran (its FFTW
    calls answered by the interpreter), the derivative, convective and pressure routines are the shims: four steps with
    two model updates vs the all-reference fixture, model state included."""
    d, meta = load("ref_full_lasd_16x16x6")
    p = O.Params(**meta["params"])
    R = shimmed(p, files=refrun.LASD_FILES, dyn_init=meta["dyn_init"], cs_count=meta["cs_count"])
    for n in ("u", "v", "w"):
        R.put(n, d[n + "0"])
    for it in range(1, max(meta["record"]) + 1):
        R.step(it, mode="full")
        if it in meta["record"]:
            for n in FIELDS:
                hi = p.nz + 1 if n in ("w", "RHSz", "p") else p.nz
                assert rel(R.get(n)[1:hi, :, :p.nx], d[f"{n}_{it}"][1:hi, :, :p.nx]) <= 1e-12, (n, it)
            for n in ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
                g = R.get(n.lower(), module="sgs_param")
                assert rel(g[1:p.nz + 1, :, :p.nx], d[f"{n}_{it}"][1:p.nz + 1, :, :p.nx]) <= (1e-10 if n == "Cs_opt2" else 1e-12), (n, it)
