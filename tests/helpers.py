"""Shared helpers for the parity tests (CUDA library on the GPU box, kernel-logic
emulator on the CPU-only container).  Test infrastructure only."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import lesgo_oracle as O  # noqa: E402
import lesgo_b200  # noqa: E402

EMUL_SO = os.path.join(ROOT, "tests", "emul", "_build", "liblesgo_emul.so")
_emul_lib = None


def emul_library():
    """Build (once) and load the CPU kernel-logic emulator: the product .cu sources compiled
    with g++ -DLESGO_EMUL.  Only tests use it; lesgo_b200 itself never loads it."""
    global _emul_lib
    if _emul_lib is None:
        subprocess.run([os.path.join(ROOT, "tests", "emul", "build_emul.sh")], check=True,
                       stdout=subprocess.DEVNULL)
        _emul_lib = lesgo_b200.Library(EMUL_SO)
    return _emul_lib


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def make_dims(p: O.Params, device=-1):
    return lesgo_b200.Dims(nx=p.nx, ny=p.ny, Nz=p.Nz, nproc=p.nproc, coord=p.coord, L_x=p.L_x, L_y=p.L_y,
                           L_z=p.L_z, lbc_mom=p.lbc_mom, ubc_mom=p.ubc_mom, sgs=p.sgs, device=device)


def random_field(p, seed, planes=None):
    rng = np.random.default_rng(seed)
    n = p.nz + 1 if planes is None else planes
    f = np.zeros((n, p.ny, p.ld))
    f[:, :, :p.nx] = rng.standard_normal((n, p.ny, p.nx))
    return f


def initial_state(p: O.Params, seed=11, amp=0.3):
    """Oracle State for rank p.coord from the synthetic global channel field."""
    u, v, w = O.synthetic_global(p.nx, p.ny, p.Nz, nproc=p.nproc, seed=seed, amp=amp, L_x=p.L_x, L_y=p.L_y,
                                 L_z=p.L_z)
    s = O.State(p)
    s.u, s.v, s.w = (O.scatter_slab(f, p) for f in (u, v, w))
    return s


# ---- the routine-by-routine comparisons, shared by the emulator and the GPU tests -----------------
def check_derivatives(core, p, tol=1e-13):
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    f = random_field(p, 1)
    out = {}
    fx, fy = O.ddxy(f, sp)
    gx, gy = core.empty(), core.empty()
    core.ddxy(f.copy(), gx, gy)
    out["ddxy_x"] = rel(gx[:, :, :nx], fx[:, :, :nx]); out["ddxy_y"] = rel(gy[:, :, :nx], fy[:, :, :nx])
    g = core.empty(); core.ddx(f.copy(), g); out["ddx"] = rel(g[:, :, :nx], fx[:, :, :nx])
    g = core.empty(); core.ddy(f.copy(), g); out["ddy"] = rel(g[:, :, :nx], fy[:, :, :nx])
    ff, fx, fy = O.filt_da(f, sp)
    h = f.copy(); core.filt_da(h, gx, gy)
    out["filt_da_f"] = rel(h[:, :, :nx], ff[:, :, :nx])
    out["filt_da_x"] = rel(gx[:, :, :nx], fx[:, :, :nx]); out["filt_da_y"] = rel(gy[:, :, :nx], fy[:, :, :nx])
    d = O.ddz_uv(f, p); g = core.empty(); core.ddz_uv(f, g)
    lo = 2 if p.coord == 0 else 1
    hi = nz - 1 if p.coord == p.nproc - 1 else nz
    out["ddz_uv"] = float(np.abs(g[lo:hi + 1, :, :nx] - d[lo:hi + 1, :, :nx]).max())
    assert g[0, 0, 0] == O.BOGUS
    d = O.ddz_w(f, p); g = core.empty(); core.ddz_w(f, g)
    lo = 1 if p.coord == 0 else 0
    out["ddz_w"] = float(np.abs(g[lo:nz, :, :nx] - d[lo:nz, :, :nx]).max())
    assert g[nz, 0, 0] == O.BOGUS
    for k, v in out.items():
        assert v <= (0.0 if k.startswith("ddz") else tol), (k, v, out)
    return out


def check_fft_raw(core, p, tol=1e-14):
    sp = O.Spectral(p)
    out = {}
    f = random_field(p, 2, planes=3)
    ref = sp.forw(f)
    got = core.fft_r2c(f.copy())
    out["r2c"] = rel(got, ref)
    back = core.fft_c2r(ref.copy())
    out["c2r"] = rel(back[:, :, :p.nx], sp.back(ref)[:, :, :p.nx])
    fb = np.zeros((2, p.ny2, p.ld_big))
    fb[:, :, :p.nx2] = np.random.default_rng(3).standard_normal((2, p.ny2, p.nx2))
    refb = sp.forw_big(fb)
    out["r2c_big"] = rel(core.fft_r2c(fb.copy(), big=True), refb)
    out["c2r_big"] = rel(core.fft_c2r(refb.copy(), big=True)[:, :, :p.nx2], sp.back_big(refb)[:, :, :p.nx2])
    # padd / unpadd are exact copies
    ub = np.full((3, p.ny2, p.ld_big), 7.0)
    core.padd(ub, ref)
    assert np.array_equal(ub, sp.padd(ref))
    cc = np.full((2, p.ny, p.ld), 7.0)
    core.unpadd(cc, refb)
    assert np.array_equal(cc, sp.unpadd(refb))
    # test_filter
    G = O.test_filter_kernel(sp)
    tf = core.test_filter(f.copy(), G)
    out["test_filter"] = rel(tf[:, :, :p.nx], O.test_filter(f, sp, G)[:, :, :p.nx])
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out


def check_convec(core, p, tol=1e-13):
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    s = O.State(p)
    for i, n in enumerate(("u", "v", "w", "dudy", "dudz", "dvdx", "dvdz", "dwdx", "dwdy")):
        setattr(s, n, random_field(p, 20 + i))
    Rx, Ry, Rz = O.convec(s, sp)
    gx, gy, gz = core.empty(), core.empty(), core.empty()
    core.convec(s.u, s.v, s.w, s.dudy, s.dudz, s.dvdx, s.dvdz, s.dwdx, s.dwdy, gx, gy, gz)
    out = {"RHSx": rel(gx[1:nz, :, :nx], Rx[1:nz, :, :nx]), "RHSy": rel(gy[1:nz, :, :nx], Ry[1:nz, :, :nx])}
    hi = nz + 1 if p.coord == p.nproc - 1 else nz
    out["RHSz"] = rel(gz[1:hi, :, :nx], Rz[1:hi, :, :nx])
    assert gx[0, 0, 0] == O.BOGUS and gx[nz, 0, 0] == O.BOGUS
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out


def check_press(core, p, tol=1e-11):
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    s = initial_state(p, seed=31)
    s.divtz = 0.1 * random_field(p, 32)
    pr, dpdx, dpdy, dpdz = O.press_stag_array(s, sp, O.LocalComm())
    g = [core.empty() for _ in range(4)]
    core.press_stag_array(s.u, s.v, s.w, s.divtz, p.dt, p.tadv1, *g)
    out = {"p": rel(g[0][0:nz + 1, :, :nx], pr[0:nz + 1, :, :nx]),
           "dpdx": rel(g[1][1:nz, :, :nx], dpdx[1:nz, :, :nx]),
           "dpdy": rel(g[2][1:nz, :, :nx], dpdy[1:nz, :, :nx]),
           "dpdz": rel(g[3][1:nz + 1, :, :nx], dpdz[1:nz + 1, :, :nx])}
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out


def step_kwargs(p, it, mode):
    return dict(dt=p.dt, tadv1=p.tadv1, tadv2=p.tadv2, first_step=(it == 0), mode=1 if mode == "full" else 0,
                ubot=p.ubot, utop=p.utop, nu=p.nu,
                mean_p_force_x=p.mean_p_force_x if p.use_mean_p_force else 0.0,
                mean_p_force_y=p.mean_p_force_y if p.use_mean_p_force else 0.0,
                sgs_model=p.sgs_model, ifilter=p.ifilter, Co=p.Co, wall_damp_exp=p.wall_damp_exp, vonk=p.vonk, zo=p.zo)


def step_kwargs_pre_dyn(p, it, mode):
    """sgs_model 5 before DYN_init: Cs_opt2 = 0.03 set at jt = 1 (sgs_stag_util.f90:187-189), no update."""
    kw = step_kwargs(p, it, mode)
    if p.sgs and p.sgs_model == 5 and mode == "full":
        kw["lasd_cs_init"] = (it == 0)
    return kw


def check_steps(core, p, nsteps=1, tol=1e-12, seed=41, names=("u", "v", "w", "p", "RHSx", "RHSy", "RHSz"), mode="core"):
    """nsteps of the device-resident step vs oracle.step(mode) from identical fields.  mode "core":
    scope rows (a)-(e); "full": + wallstress, constant-coefficient sgs_stag, divstress (rows (f)-1)."""
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    G = O.test_filter_kernel(sp)
    s = initial_state(p, seed=seed)
    for n in ("u", "v", "w"):
        core.upload(n, getattr(s, n))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(core.dims.shape))
    for it in range(nsteps):
        O.step(s, sp, O.LocalComm(), mode=mode, first_step=(it == 0), G_test=G)
        core.step(**step_kwargs_pre_dyn(p, it, mode))
    out = {}
    for n in names:
        g = core.download(n)
        r = getattr(s, n)
        hi = nz + 1 if n in ("w", "RHSz", "p") else nz
        out[n] = rel(g[1:hi, :, :nx], r[1:hi, :, :nx])
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out


def lasd_schedule(p, it, cs_count=2, dyn_init=2):
    """Step counters of sgs_stag_util.f90:73-82,183-216 for a fresh run (inilag, jt = jt_total = it + 1,
    fixed dt): Cs_opt2 = 0.03 at jt = 1, lagrange_Sdep every cs_count steps from DYN_init on, F_* initialised
    the first time (jt == DYN_init)."""
    jt = it + 1
    return dict(lasd_cs_init=(jt == 1), lasd_update=(jt >= dyn_init and jt % cs_count == 0),
                lasd_init_F=(jt == dyn_init), lagran_dt=cs_count * p.dt)


class LasdClock:
    """The step counters AND the Lagrangian time interval of sgs_stag_util.f90:73-82,183-216 / lagrange_Sdep.f90:266-267,
    430 for a fresh run (inilag; jt = jt_total), as fortran/lesgo_gpu_resident_mod.f90: gpu_lasd_switches keeps them for
    lesgo_gpu_step: with use_cfl_dt lagran_dt is ACCUMULATED (+ dt on every step from jt = DYN_init - cs_count + 1 on)
    and zeroed by the step that runs lagrange_Sdep; otherwise it is cs_count * dt."""

    def __init__(self, cs_count, dyn_init, use_cfl_dt):
        self.cs_count, self.dyn_init, self.use_cfl_dt = cs_count, dyn_init, use_cfl_dt
        self.acc, self.initialised = 0.0, False

    def switches(self, jt, dt):
        if self.use_cfl_dt:
            if jt >= self.dyn_init - self.cs_count + 1:
                self.acc += dt
            lagran_dt = self.acc
        else:
            lagran_dt = self.cs_count * dt
        out = dict(lasd_cs_init=(jt == 1), lasd_update=False, lasd_init_F=False, lagran_dt=lagran_dt)
        if jt != 1 and jt >= self.dyn_init and jt % self.cs_count == 0:
            out["lasd_update"] = True
            self.acc = 0.0
            if not self.initialised and (jt == self.cs_count or jt == self.dyn_init):
                out["lasd_init_F"] = True
                self.initialised = True
        return out


def check_lasd_steps(core, p, nsteps=4, tol=1e-11, seed=61, amp=0.5):
    """Full steps with sgs_model = 5 (Lagrangian scale-dependent dynamic model, rows (f)-2) against the oracle:
    velocities, pressure, and the model's own state F_LM, F_MM, F_QN, F_NN, Cs_opt2."""
    assert p.sgs and p.sgs_model == 5
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    G = O.test_filter_kernel(sp)
    G2 = O.test_filter_kernel(sp, alpha=4.0)
    s = initial_state(p, seed=seed, amp=amp)
    for n in ("u", "v", "w"):
        core.upload(n, getattr(s, n))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz", "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
        core.upload(n, np.zeros(core.dims.shape))
    for it in range(nsteps):
        sch = lasd_schedule(p, it)
        lasd = dict(sp=sp, G_test=G, G_test_test=G2, lagran_dt=sch["lagran_dt"], cs_init=sch["lasd_cs_init"],
                    update=sch["lasd_update"], init_F=sch["lasd_init_F"])
        O.step(s, sp, O.LocalComm(), mode="full", first_step=(it == 0), G_test=G, lasd=lasd)
        core.step(**step_kwargs(p, it, "full"), **sch)
    out = {}
    for n in ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz", "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
        g = core.download(n)
        r = getattr(s, n)
        hi = nz + 1 if n in ("w", "RHSz", "p", "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2") else nz
        out[n] = rel(g[1:hi, :, :nx], r[1:hi, :, :nx])
    for k, v in out.items():
        # Cs_opt2 = F_LM / F_MM / clip(Cs_4d / Cs_2d) is a ratio of ratios of running averages: locally
        # ill-conditioned where F_LM -> 0, so its own bound is looser than that of the fields it feeds
        assert v <= (CS_TOL if k == "Cs_opt2" else tol), (k, v, out)
    return out


CS_TOL = 1e-9


def make_farm(p, comm=None, tilt=True, overlap=False):
    """Two actuator disks (one yawed and tilted, so all three force components are exercised) on the
    grid of p, with the node lists turbines_nodes (turbines.f90:275-462) builds."""
    comm = comm or O.LocalComm()
    dia = 0.3 * p.L_y
    h = 0.45 * p.L_z
    dlt = 1.5 * np.sqrt(p.dx ** 2 + p.dy ** 2 + p.dz ** 2)
    farm = [O.Turbine(xloc=0.25 * p.L_x, yloc=0.3 * p.L_y, height=h, dia=dia, thk=1.2 * p.dx, theta1=0.0, theta2=0.0,
                      Ct_prime=1.33, u_d_T=-0.7),
            O.Turbine(xloc=0.97 * p.L_x, yloc=0.8 * p.L_y, height=0.9 * h, dia=0.8 * dia, thk=1.2 * p.dx,
                      theta1=20.0 if tilt else 0.0, theta2=10.0 if tilt else 0.0, Ct_prime=1.0, u_d_T=-0.5)]
    if overlap:
        # a third disk that shares grid points with the first: the reference assigns node by node in disk
        # order (turbines.f90:599-606), so the later disk's force replaces the earlier one's there
        farm.append(O.Turbine(xloc=0.25 * p.L_x + 1.0 * p.dx, yloc=0.3 * p.L_y + 2.0 * p.dy, height=h, dia=0.7 * dia,
                              thk=1.2 * p.dx, Ct_prime=0.8, u_d_T=-0.4))
    for t in farm:
        val = O.standin_indicator(t.dia, t.thk, dlt, dlt)
        O.turbines_nodes(p, [t], val, comm)
    return farm


def check_turbines(core, p, nsteps=2, tol=1e-12, mode="core", eps=0.3, adm_correction=True, overlap=False, rotation=None):
    """turbines_forcing standalone (force fields, per-disk scalars) and inside lesgo_gpu_step.
    rotation = tip speed ratio: the ADM with rotation (use_rotation, turbines.f90:607-615)."""
    rkw = dict(use_rotation=True, tip_speed_ratio=float(rotation)) if rotation else {}
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    G = O.test_filter_kernel(sp)
    s = initial_state(p, seed=71)
    s.u += 1.0
    farm = make_farm(p, overlap=overlap)
    assert all(len(t.ind) > 10 for t in farm)
    if overlap:
        shared = set(map(tuple, farm[0].nodes)) & set(map(tuple, farm[2].nodes))
        assert len(shared) > 5, "the overlap case needs shared grid points"
    for n in ("u", "v", "w"):
        core.upload(n, getattr(s, n))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(core.dims.shape))
    core.turbines_init(farm, adm_correction=adm_correction, **rkw)
    out = {}
    # standalone call; the oracle copy keeps the running averages of the two sides in step
    import copy
    farm0 = copy.deepcopy(farm)
    fx, fy, fz = O.turbines_forcing(s, p, O.LocalComm(), farm0, eps, adm_correction=adm_correction, **rkw)
    u_d, u_d_T, f_n = core.turbines_forcing(eps)
    out["u_d"] = rel(u_d, [t.u_d for t in farm0]); out["u_d_T"] = rel(u_d_T, [t.u_d_T for t in farm0])
    out["f_n"] = rel(f_n, [t.f_n for t in farm0])
    for n, r in (("fxa", fx), ("fya", fy), ("fza", fz)):
        g = core.download(n)
        assert np.count_nonzero(r[1:nz, :, :nx]) > 0, n
        out[n] = rel(g[1:nz, :, :nx], r[1:nz, :, :nx])
    core.turbines_init(farm, adm_correction=adm_correction, **rkw)      # reset the running averages
    for it in range(nsteps):
        O.step(s, sp, O.LocalComm(), mode=mode, first_step=(it == 0), G_test=G,
               turbines=dict(farm=farm, eps=eps, adm_correction=adm_correction, **rkw))
        core.step(**step_kwargs_pre_dyn(p, it, mode), turbines=True, turbines_eps=eps)
    for n in ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz"):
        g = core.download(n)
        r = getattr(s, n)
        hi = nz + 1 if n in ("w", "RHSz", "p") else nz
        out["step_" + n] = rel(g[1:hi, :, :nx], r[1:hi, :, :nx])
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out


TAVG_SEAM = ("w_uv", "u_w", "v_w", "uw", "p", "vortz", "fz", "u2")


def farm_for_rank(farm, p):
    """The slab of rank p.coord of a single-slab farm: nodes with global k in the rank's 1..nz-1 (turbines.f90:246-247,
    425), same normalised weights."""
    import copy
    lo, hi = 1 + p.coord * (p.nz - 1), (p.nz - 1) * (p.coord + 1)
    out = []
    for t in farm:
        c = copy.copy(t)
        m = (t.nodes[:, 2] >= lo) & (t.nodes[:, 2] <= hi)
        c.nodes = t.nodes[m].copy()
        c.nodes[:, 2] -= p.coord * (p.nz - 1)
        c.ind = t.ind[m].copy()
        if t.ind_t is not None:
            c.ind_t, c.e_theta = t.ind_t[m].copy(), t.e_theta[m].copy()
        out.append(c)
    return out


def check_multirank_steps(lib, kw, nproc, nsteps=2, tol=1e-11, seed=51, device_of=None, mode="core", lasd=False,
                          turbines=False, tavg=False, p2p=False, local=False, ref_global=None, rotation=None, ref_tavg=None,
                          farm=None, adm_correction=False, fields0=None, ref_disks=None):
    """nproc ranks (threads of this process, one Core each) advance `nsteps` core steps;
    the gathered result must match the SINGLE-slab oracle (which the multi-slab oracle
    equals, tests/test_oracle_kat.py)."""
    import threading
    pg = O.Params(nproc=1, **kw)
    spg = O.Spectral(pg)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)
    if fields0 is not None:
        ug, vg, wg = fields0
    sref = O.State(pg)
    sref.u, sref.v, sref.w = (O.scatter_slab(f, pg) for f in (ug, vg, wg))
    Gg = O.test_filter_kernel(spg)
    G2g = O.test_filter_kernel(spg, alpha=4.0)
    names = ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz") + (("Cs_opt2", "F_LM", "F_NN") if lasd else ())
    farm_g = (farm if farm is not None else make_farm(pg)) if turbines else None
    farm_ref = [__import__("copy").copy(t) for t in farm_g] if turbines else None
    tkw = dict(turbines=True, turbines_eps=0.3) if turbines else {}
    rkw = dict(use_rotation=True, tip_speed_ratio=float(rotation)) if rotation else {}    # turbines.f90:607-615
    for it in range(nsteps):
        ld_ = None
        if lasd:
            sch = lasd_schedule(pg, it)
            ld_ = dict(sp=spg, G_test=Gg, G_test_test=G2g, lagran_dt=sch["lagran_dt"], cs_init=sch["lasd_cs_init"],
                       update=sch["lasd_update"], init_F=sch["lasd_init_F"])
        O.step(sref, spg, O.LocalComm(), mode=mode, first_step=(it == 0), G_test=Gg, lasd=ld_,
               turbines=dict(farm=farm_ref, eps=0.3, adm_correction=adm_correction, **rkw) if turbines else None)
        if tavg:
            if it == 0:
                tref = O.Tavg(pg)
            O.tavg_compute(tref, sref, pg, O.LocalComm(), pg.dt, forces=turbines)
    ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
    cores = [lesgo_b200.Core(make_dims(p, device=(device_of(p.coord) if device_of else -1)), lib=lib) for p in ps]
    ident = cores[0].comm_unique_id(local=local)      # local: every rank on ONE GPU (single-device transport)
    res, err = [None] * nproc, [None] * nproc
    blobs, bar = [None] * nproc, threading.Barrier(nproc)

    def work(r):
        try:
            p, c = ps[r], cores[r]
            c.comm_init(ident)
            if p2p:     # pressure transposes over peer memory instead of NCCL all-to-alls
                blobs[r] = c.comm_p2p_export()
                bar.wait(timeout=600)
                c.comm_p2p_import(blobs)
            for n, g in (("u", ug), ("v", vg), ("w", wg)):
                c.upload(n, O.scatter_slab(g, p))
            for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz") + (("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2") if lasd else ()):
                c.upload(n, np.zeros(c.dims.shape))
            if turbines:
                c.turbines_init(farm_for_rank(farm_g, p), adm_correction=adm_correction, **rkw)
            for it in range(nsteps):
                c.step(**step_kwargs(p, it, mode), **(lasd_schedule(p, it) if lasd else {}), **tkw)
                if tavg:
                    c.tavg_compute(p.dt)
            res[r] = {n: c.download(n) for n in names}
            if tavg:
                for n in TAVG_SEAM:
                    res[r]["tavg_" + n] = c.tavg_download(n)[0]
            if turbines:
                res[r]["u_d_T"] = c.turbines_forcing(0.3)[1]
            # mpi_sync_real_array (mpi_defs.f90:245-262) on a host array
            var = np.zeros(c.dims.shape)
            for k in range(p.nz + 1):
                var[k] = 100.0 * r + k
            c.sync_real_array(var, 3)
            if r > 0:
                assert np.all(var[0] == 100.0 * (r - 1) + p.nz - 1), "SYNC_UP"
            else:
                assert np.all(var[0] == 0.0)
            if r < nproc - 1:
                assert np.all(var[p.nz] == 100.0 * (r + 1) + 1), "SYNC_DOWN"
            else:
                assert np.all(var[p.nz] == 100.0 * r + p.nz)
            res[r]["cfl"] = c.max_cfl(p.dt)
            res[r]["cfl_dt"] = c.cfl_dt(0.0625)
        except BaseException as e:  # noqa
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nproc)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for e in err:
        if e is not None:
            raise e
    out = {}
    nzt = pg.nz_tot
    for n in names:
        top = n in ("w", "RHSz", "p", "Cs_opt2", "F_LM", "F_NN")
        g = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=top)
        hi = nzt if top else nzt - 1
        out[n] = rel(g[1:hi + 1, :, :pg.nx], getattr(sref, n)[1:hi + 1, :, :pg.nx])
        if ref_global is not None:         # ... and against fields the reference's own MPI code path produced
            out["ref_" + n] = rel(g[1:hi + 1, :, :pg.nx], ref_global[n][1:hi + 1, :, :pg.nx])
    if tavg:
        # the accumulators whose interpolations cross the slab seams
        for n in TAVG_SEAM:
            g = O.gather_slabs([res[r]["tavg_" + n] for r in range(nproc)], ps, top_extra=False)
            out["tavg_" + n] = rel(g[1:nzt], getattr(tref, n)[1:nzt])
            if ref_tavg is not None:       # ... and against the accumulators of the reference's own MPI run
                out["ref_tavg_" + n] = rel(g[1:nzt], ref_tavg[n][1:nzt])
    if turbines:
        # one more forcing call on both sides: every rank must hold the same (global) disk velocities
        if ref_disks is not None:      # the disks' running averages after nsteps against the reference's own MPI run
            out["ref_u_d_T"] = rel([t.u_d_T for t in farm_ref], ref_disks)
        O.turbines_forcing(sref, pg, O.LocalComm(), farm_ref, 0.3, adm_correction=adm_correction, **rkw)
        for r in range(nproc):
            out[f"u_d_T_rank{r}"] = rel(res[r]["u_d_T"], [t.u_d_T for t in farm_ref])
    cfl_ref = O.get_max_cfl(sref, pg, O.LocalComm())
    assert all(abs(res[r]["cfl"] - cfl_ref) <= 1e-12 * cfl_ref for r in range(nproc)), (cfl_ref, [res[r]["cfl"] for r in range(nproc)])
    dt_ref = O.get_cfl_dt(sref, pg, O.LocalComm(), 0.0625)
    assert all(abs(res[r]["cfl_dt"] - dt_ref) <= 1e-12 * dt_ref for r in range(nproc)), (dt_ref, [res[r]["cfl_dt"] for r in range(nproc)])
    for k, v in out.items():
        assert v <= (CS_TOL if k == "Cs_opt2" else tol), (k, v, out)
    return out


def check_tavg(core, p, nsteps=2, tol=1e-12, turbines=False):
    """tavg%compute (time_average.f90:176-320) after each of nsteps full steps: all 26 accumulators."""
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    G = O.test_filter_kernel(sp)
    s = initial_state(p, seed=81)
    s.u += 1.0
    O.lasd_alloc(s)
    s.Cs_opt2[...] = 0.01 * np.abs(random_field(p, 82))
    for n in ("u", "v", "w", "Cs_opt2"):
        core.upload(n, getattr(s, n))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(core.dims.shape))
    farm = make_farm(p) if turbines else None
    if turbines:
        core.turbines_init(farm)
    t = O.Tavg(p)
    for it in range(nsteps):
        O.step(s, sp, O.LocalComm(), mode="full", first_step=(it == 0), G_test=G,
               turbines=dict(farm=farm, eps=0.3) if turbines else None)
        core.step(**step_kwargs(p, it, "full"), **(dict(turbines=True, turbines_eps=0.3) if turbines else {}))
        dt_avg = p.dt * (1 + it)                                   # variable increments, as with use_cfl_dt
        O.tavg_compute(t, s, p, O.LocalComm(), dt_avg, forces=turbines)
        core.tavg_compute(dt_avg)
    out = {}
    for n in O.TAVG_FIELDS:
        g, tt = core.tavg_download(n)
        r = getattr(t, n)
        out[n] = rel(g[1:nz], r[1:nz]) if np.any(r[1:nz]) else float(np.abs(g[1:nz]).max())
        assert abs(tt - t.total_time) < 1e-15
    assert turbines or out["fx"] == 0.0
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    core.tavg_reset()
    g, tt = core.tavg_download("u2")
    assert tt == 0.0 and not g.any()
    return out


def check_checkpoint(make_core, p, tmp_path, monkeypatch):
    """Restart file: device -> file -> oracle reader and scipy's independent Fortran-record reader; oracle writer ->
    device; gfortran subrecords (forced small) round trip; wrong-grid files are refused."""
    import scipy.io
    s = O.State(p)
    O.lasd_alloc(s)
    for i, n in enumerate(O.CHECKPOINT_FIELDS):
        getattr(s, n)[...] = random_field(p, 300 + i)
    core = make_core()
    for n in O.CHECKPOINT_FIELDS:
        core.upload(n, getattr(s, n))
    f1 = str(tmp_path / "vel.out.c0")
    core.checkpoint_write(f1)
    s2 = O.State(p)
    O.checkpoint_read(s2, p, f1)
    for n in O.CHECKPOINT_FIELDS:
        assert np.array_equal(getattr(s2, n)[1:], getattr(s, n)[1:]), n
    rec = scipy.io.FortranFile(f1, "r").read_reals(np.float64)
    assert np.array_equal(rec.reshape(11, p.nz, p.ny, p.ld)[3], s.RHSx[1:])
    # oracle-written file -> device
    f2 = str(tmp_path / "vel.in.c0")
    O.checkpoint_write(s, p, f2)
    assert open(f1, "rb").read() == open(f2, "rb").read()
    core2 = make_core()
    core2.checkpoint_read(f2)
    for n in O.CHECKPOINT_FIELDS:
        assert np.array_equal(core2.download(n)[1:], getattr(s, n)[1:]), n
    # subrecords: limit of 1000 bytes -> many subrecords with signed markers
    monkeypatch.setenv("LESGO_SUBRECORD_MAX", "1000")
    f3 = str(tmp_path / "vel.sub.c0")
    core.checkpoint_write(f3)
    raw = np.fromfile(f3, dtype=np.uint8)
    nsub = -(-(11 * p.nz * p.ny * p.ld * 8) // 1000)
    assert raw.size == 11 * p.nz * p.ny * p.ld * 8 + 8 * nsub
    assert raw[:4].view(np.int32)[0] == -1000 and raw[1004:1008].view(np.int32)[0] == 1000      # first: more follow / none before
    assert raw[1008:1012].view(np.int32)[0] == -1000 and raw[2012:2016].view(np.int32)[0] == -1000
    core3 = make_core()
    core3.checkpoint_read(f3)
    for n in O.CHECKPOINT_FIELDS:
        assert np.array_equal(core3.download(n)[1:], getattr(s, n)[1:]), n
    # subrecord boundaries at awkward places: one byte short of the record, exactly the record, 8 bytes, a prime
    total = 11 * p.nz * p.ny * p.ld * 8
    for lim in (total - 1, total, total + 1, 8, 4099):
        monkeypatch.setenv("LESGO_SUBRECORD_MAX", str(lim))
        fx = str(tmp_path / f"vel.sub{lim}.c0")
        core.checkpoint_write(fx)
        assert os.path.getsize(fx) == total + 8 * (-(-total // lim)), lim
        cx = make_core()
        cx.checkpoint_read(fx)
        for n in ("u", "F_NN"):
            assert np.array_equal(cx.download(n)[1:], getattr(s, n)[1:]), (lim, n)
    monkeypatch.delenv("LESGO_SUBRECORD_MAX")
    with pytest_raises_library("record length"):
        core3.checkpoint_read(f3)                               # written with other markers than expected now
    with pytest_raises_library("cannot open"):
        core3.checkpoint_read(str(tmp_path / "missing"))
    return True


def pytest_raises_library(match):
    import pytest
    return pytest.raises(lesgo_b200.LibraryError, match=match)


def check_misc(core, p):
    """wavenumbers (fft.f90:130-160), tridag_array with caller-supplied coefficients
    (tridag_array.f90:166-246) and its zero-pivot error path."""
    sp = O.Spectral(p)
    kx, ky, k2 = core.wavenumbers()
    assert np.array_equal(kx, sp.kx) and np.array_equal(ky, sp.ky) and np.array_equal(k2, sp.k2)
    n = p.nz + 1
    rng = np.random.default_rng(77)
    a = rng.uniform(0.5, 1.0, (n, p.ny, p.lh)); c = rng.uniform(0.5, 1.0, (n, p.ny, p.lh))
    b = -(a + c + rng.uniform(0.1, 1.0, (n, p.ny, p.lh)))
    r = np.zeros((n, p.ny, p.ld)); r[:, :, :] = rng.standard_normal((n, p.ny, p.ld))
    u = np.zeros_like(r)
    core.tridag_array(a, b, c, r, u)
    # oracle layout: row j at index j (index 0 unused), n rows = nz + 1
    pad = lambda x: np.concatenate([np.zeros((1,) + x.shape[1:]), x])
    uo = pad(np.zeros_like(r))
    po = O.Params(nx=p.nx, ny=p.ny, Nz=p.Nz, nproc=1)
    O.tridag_array(pad(a), pad(b), pad(c), pad(r), uo, po, O.LocalComm())
    M = np.zeros((p.ny, p.lh), bool); M[:, :p.lh - 1] = True; M[p.ny // 2] = False; M[0, 0] = False
    got = u.view(np.complex128)[:, M]; ref = uo[1:].view(np.complex128)[:, M]
    assert rel(got, ref) < 1e-14, rel(got, ref)
    b0 = b.copy(); b0[0, 1, 1] = 0.0
    try:
        core.tridag_array(a, b0, c, r, u)
        raise AssertionError("zero pivot not reported")
    except lesgo_b200.LibraryError as e:
        assert "zero pivot" in str(e)
    return True


def check_fftw_shim(lib, core, p, big_generic=None):
    """The FFTW3 legacy-Fortran symbols the library exports (include/lesgo_gpu.h): called by
    reference exactly as gfortran calls them -- fft.f90:114-121 (in-place plans of module fft, seven
    arguments), turbine_indicator.f90:130-151 (out-of-place plans of another size, destroyed after use)."""
    import ctypes as C
    sp = O.Spectral(p)
    out = {}
    byref = C.byref
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    flags = C.c_int(32)                                     # FFTW_PATIENT; ignored

    def plan(kind, n0, n1, a, b):
        h = C.c_longlong(-7)
        fn = lib.dfftw_plan_dft_r2c_2d if kind == "r2c" else lib.dfftw_plan_dft_c2r_2d
        fn(byref(h), byref(C.c_int(n0)), byref(C.c_int(n1)), ptr(a), ptr(b), byref(flags))
        assert h.value > 0
        return h

    # module fft's four plans: in place on (ld, ny) / (ld_big, ny2) arrays
    data = np.zeros((p.ny, p.ld)); data_big = np.zeros((p.ny2, p.ld_big))
    forw, back = plan("r2c", p.nx, p.ny, data, data), plan("c2r", p.nx, p.ny, data, data)
    forw_big, back_big = plan("r2c", p.nx2, p.ny2, data_big, data_big), plan("c2r", p.nx2, p.ny2, data_big, data_big)
    assert len({forw.value, back.value, forw_big.value, back_big.value}) == 4
    f = random_field(p, 5, planes=1)
    ref = sp.forw(f)
    g = f[0].copy()
    lib.dfftw_execute_dft_r2c(byref(forw), ptr(g), ptr(g))
    out["forw"] = rel(g, ref[0])
    h = ref[0].copy()
    lib.dfftw_execute_dft_c2r(byref(back), ptr(h), ptr(h))
    out["back"] = rel(h[:, :p.nx], sp.back(ref)[0][:, :p.nx])
    fb = np.zeros((1, p.ny2, p.ld_big)); fb[:, :, :p.nx2] = np.random.default_rng(6).standard_normal((1, p.ny2, p.nx2))
    refb = sp.forw_big(fb)
    g = fb[0].copy()
    lib.dfftw_execute_dft_r2c(byref(forw_big), ptr(g), ptr(g))
    out["forw_big"] = rel(g, refb[0])
    h = refb[0].copy()
    lib.dfftw_execute_dft_c2r(byref(back_big), ptr(h), ptr(h))
    out["back_big"] = rel(h[:, :p.nx2], sp.back_big(refb)[0][:, :p.nx2])

    # plans of another shape, out of place (turbine_indicator.f90: g(N,N) -> ghat(N/2+1,N), fxhat -> fx)
    for n0, n1 in ((40, 24), (30, 50)) + ((big_generic,) if big_generic else ()):
        rng = np.random.default_rng(n0 + n1)
        x = rng.standard_normal((n1, n0))
        xh = np.zeros((n1, n0 // 2 + 1), complex)
        pl = plan("r2c", n0, n1, x, xh)
        lib.dfftw_execute_dft_r2c(byref(pl), ptr(x), ptr(xh))
        lib.dfftw_destroy_plan(byref(pl))
        assert pl.value == 0
        out[f"gen_r2c_{n0}x{n1}"] = rel(xh, np.fft.rfft2(x))
        # c2r of a spectrum that is NOT Hermitian in its kx = 0 / Nyquist columns: FFTW transforms y first
        # (complex) and x last (c2r, imaginary parts of the two self-conjugate entries ignored)
        yh = rng.standard_normal((n1, n0 // 2 + 1)) + 1j * rng.standard_normal((n1, n0 // 2 + 1))
        y = np.zeros((n1, n0))
        pl = plan("c2r", n0, n1, yh, y)
        keep = yh.copy()
        lib.dfftw_execute_dft_c2r(byref(pl), ptr(yh), ptr(y))
        lib.dfftw_destroy_plan(byref(pl))
        want = np.fft.irfft(np.fft.ifft(keep, axis=0), n=n0, axis=1) * (n0 * n1)
        out[f"gen_c2r_{n0}x{n1}"] = rel(y, want)
    # generic IN-PLACE plan (padded real rows), shape different from the context's
    n0, n1 = 20, 12
    z = np.zeros((n1, n0 + 2)); z[:, :n0] = np.random.default_rng(9).standard_normal((n1, n0))
    want = np.fft.rfft2(z[:, :n0])
    pl = plan("r2c", n0, n1, z, z)
    lib.dfftw_execute_dft_r2c(byref(pl), ptr(z), ptr(z))
    out["gen_inplace"] = rel(z.view(complex), want)
    # error paths of the status-returning forms: executing out of place with an in-place plan, a destroyed plan,
    # lengths that are not 2-3-5 smooth
    w = np.zeros_like(z)
    assert lib.fftw_execute(pl.value, 0, ptr(z), ptr(w)) != 0 and b"in place" in lib.fftw_last_error()
    assert lib.fftw_execute(pl.value, 1, ptr(z), ptr(z)) != 0
    hv = pl.value
    lib.dfftw_destroy_plan(byref(pl))
    assert lib.fftw_execute(hv, 0, ptr(z), ptr(z)) != 0 and b"plan handle" in lib.fftw_last_error()
    bad = C.c_longlong(0)
    assert lib.fftw_plan_2d(0, 14, 16, 1, byref(bad)) != 0 and bad.value == 0
    for k, v in out.items():
        assert v <= 1e-13, (k, v, out)
    return out


def check_cfl(core, p, seed=91):
    """get_max_cfl / get_cfl_dt (cfl_util.f90:35-113) on one slab, resident fields."""
    s = initial_state(p, seed=seed)
    for n in ("u", "v", "w"):
        core.upload(n, getattr(s, n))
    c_ref = O.get_max_cfl(s, p, O.LocalComm())
    d_ref = O.get_cfl_dt(s, p, O.LocalComm(), 0.0625)
    c_got, d_got = core.max_cfl(p.dt), core.cfl_dt(0.0625)
    out = {"max_cfl": abs(c_got - c_ref) / c_ref, "cfl_dt": abs(d_got - d_ref) / d_ref}
    for k, v in out.items():
        assert v <= 1e-15, (k, v, out)
    # the maximum sits in one component: make each of u, v, w the binding one in turn
    for n in ("v", "w"):
        t = getattr(s, n).copy(); t[p.nz // 2, 3, 5] = 50.0
        core.upload(n, t)
        setattr(s, n, t)
        assert abs(core.cfl_dt(0.1) - O.get_cfl_dt(s, p, O.LocalComm(), 0.1)) <= 1e-15 * core.cfl_dt(0.1), n
    return out


def check_variable_dt_steps(core, p, nsteps=10, cfl=0.0625, tol=1e-9, mode="core", seed=43):
    """use_cfl_dt = .true. as LES_channel_Re1000 ships it (lesgo.conf:117): every step takes dt = get_cfl_dt(),
    tadv1 = 1 + dt/(2 dt_f), tadv2 = 1 - tadv1 (main.f90:135-144); the first step is forced to first-order Euler by
    dt_f = get_cfl_dt * huge (initialize.f90:192-199) and RHS_f = RHS (main.f90:273-280).  Each side computes its OWN
    dt from its own fields (device: lesgo_gpu_cfl_dt), so the dt sequences are compared too."""
    import dataclasses
    p = dataclasses.replace(p)                     # the oracle driver mutates dt, tadv1, tadv2
    sp = O.Spectral(p)
    nx, nz = p.nx, p.nz
    G = O.test_filter_kernel(sp)
    s = initial_state(p, seed=seed)
    for n in ("u", "v", "w"):
        core.upload(n, getattr(s, n))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(core.dims.shape))
    O.cfl_dt_start(s, p, O.LocalComm(), cfl)
    import sys as _sys
    dt_dev = core.cfl_dt(cfl) * _sys.float_info.max
    dts = []
    for it in range(nsteps):
        O.cfl_dt_advance(s, p, O.LocalComm(), cfl)
        O.step(s, sp, O.LocalComm(), mode=mode, first_step=(it == 0), G_test=G)
        dt_f = dt_dev
        dt_dev = core.cfl_dt(cfl)
        t1 = 1.0 + 0.5 * dt_dev / dt_f
        kw = step_kwargs_pre_dyn(p, it, mode)
        kw.update(dt=dt_dev, tadv1=t1, tadv2=1.0 - t1)
        core.step(**kw)
        dts.append((dt_dev, p.dt, t1, p.tadv1))
    assert dts[0][2] == 1.0 and dts[0][3] == 1.0, "Euler start"
    assert nsteps == 1 or len({d[0] for d in dts}) > 1, "dt must actually vary"
    out = {"dt": max(abs(a - b) / b for a, b, _, _ in dts), "tadv1": max(abs(a - b) for _, _, a, b in dts)}
    for n in ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz"):
        g = core.download(n)
        r = getattr(s, n)
        hi = nz + 1 if n in ("w", "RHSz", "p") else nz
        out[n] = rel(g[1:hi, :, :nx], r[1:hi, :, :nx])
    for k, v in out.items():
        assert v <= tol, (k, v, out)
    return out
