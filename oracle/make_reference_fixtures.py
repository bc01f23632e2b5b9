#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz: outputs of the REFERENCE'S OWN SOURCES (under /root/reference), executed by
oracle/f90exec.py through oracle/refrun.py.  Run in the development container (the GPU boxes have no
/root/reference):

    python oracle/make_reference_fixtures.py

Each fixture stores the Params it was made with, the initial u, v, w, and the reference's fields after the
recorded steps; tests/test_reference_pin.py compares the oracle (CPU suite) and the CUDA path (-m gpu) with them.
The FFTs inside the reference run are pocketfft (FFTW3 is not available anywhere), everything else is
the reference's text: index ranges, wall-plane cases, BOGUS planes, operation order, the time-loop glue of
main.f90, project, cfl_util.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import lesgo_oracle as O          # noqa: E402
from oracle import refrun                      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STEP_FIELDS = ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz")

STEP_CASES = {
    # BASELINE configs[1] in miniature: DNS Couette walls, core path only (rows a-e)
    "ref_core_couette_32x32x8": dict(kw=dict(nx=32, ny=32, Nz=8, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, L_x=4 * np.pi),
                                     mode="core", record=(1, 10)),
    # the same with the molecular stress: wallstress (DNS), calc_Sij, sgs_stag (sgs = .false.), divstress_uv/w
    "ref_full_couette_16x16x8": dict(kw=dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, sgs=False, molec=True,
                                             nu_molec=1e-2), mode="full", record=(1, 10)),
    # LES half channel as LES_channel_Re1000: equilibrium wall model below, stress-free lid, Smagorinsky + Mason damping
    "ref_full_les_channel_16x32x8": dict(kw=dict(nx=16, ny=32, Nz=8, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False,
                                                 use_mean_p_force=True, mean_p_force_x=1.0, L_y=np.pi), mode="full", record=(1, 10)),
    # stress-free on both sides, LES branch of convec (jzLo = 2), core path
    "ref_core_free_slip_les_16x16x6": dict(kw=dict(nx=16, ny=16, Nz=6, lbc_mom=0, ubc_mom=0, sgs=True, L_x=4.0, L_y=3.0),
                                           mode="core", record=(1, 4)),
    # use_cfl_dt as the shipped LES_channel_Re1000 (lesgo.conf:117): sgs_model 5 before DYN_init, variable dt
    "ref_full_cfl_dt_sgs5_16x16x8": dict(kw=dict(nx=16, ny=16, Nz=8, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=5, molec=False),
                                         mode="full", record=(1, 10), cfl=0.0625),
}


LASD_FIELDS = ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2")


def run_lasd_case(name="ref_full_lasd_16x16x6", nsteps=4):
    """Rows (f)-2 from the reference text: lagrange_Sdep.f90 + interpolag_Sdep.f90 + trilinear_interp_w / cell_indx_w
    (functions.f90) + grid_m, inside full steps with sgs_model = 5, DYN_init = cs_count = 2 (Cs_opt2 = 0.03 at jt = 1,
    updates at jt = 2 and 4, F_* initialised at jt = 2)."""
    kw = dict(nx=16, ny=16, Nz=6, L_x=4.0, L_y=3.0, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False, dt=2e-3)
    p = O.Params(**kw)
    R = refrun.Reference(p, files=refrun.LASD_FILES, dyn_init=2, cs_count=2)
    u, v, w = initial_fields(p, seed=61, amp=0.5)
    out = {"u0": u, "v0": v, "w0": w}
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    t0 = time.time()
    for it in range(1, nsteps + 1):
        R.step(it, mode="full")
        if it in (2, nsteps):
            for n in STEP_FIELDS:
                out[f"{n}_{it}"] = R.get(n)
            for n in LASD_FIELDS:
                out[f"{n}_{it}"] = R.get(n.lower(), module="sgs_param")
    meta = dict(params=params_record(p), mode="full", record=[2, nsteps], cfl=None, dyn_init=2, cs_count=2,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py",
                statements=R.I.nstmt)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: {nsteps} steps, {R.I.nstmt} reference statements, {time.time() - t0:.1f} s")


def run_lasd_cfl_case(name="ref_full_lasd_cfl_dt_16x16x6", nsteps=6, cfl=0.0625, dyn_init=2, cs_count=2):
    """Rows (f)-2 with use_cfl_dt as the shipped lesgo.conf runs it: the time step varies, so lagran_dt is ACCUMULATED
    (sgs_stag_util.f90:73-82: + dt on every step from jt = DYN_init - cs_count + 1 on, zeroed by lagrange_Sdep.f90:430)
    instead of being cs_count * dt.  DYN_init = cs_count = 2, six steps: updates at jt = 2 (lagran_dt = dt1 + dt2, F_*
    initialised), jt = 4 (dt3 + dt4) and jt = 6 (dt5 + dt6)."""
    kw = dict(nx=16, ny=16, Nz=6, L_x=4.0, L_y=3.0, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False)
    p = O.Params(**kw)
    R = refrun.Reference(p, files=refrun.LASD_FILES, dyn_init=dyn_init, cs_count=cs_count)
    u, v, w = initial_fields(p, seed=63, amp=0.5)
    out = {"u0": u, "v0": v, "w0": w}
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    R.I.set("param", "cfl", float(cfl)); R.I.set("param", "use_cfl_dt", True); R.I.set("param", "cfl_f", 0.0)
    R.I.exec_lines(os.path.join(refrun.REF, "initialize.f90"), 192, 200, ["types", "param", "sim_param", "cfl_util"],
                   local={"dt_dim": 0.0})
    dts = []
    record = (dyn_init, nsteps)
    for it in range(1, nsteps + 1):
        R.cfl_dt_step_setup()                                      # main.f90:135-144
        dts.append((R.I.get("param", "dt"), R.I.get("param", "tadv1"), R.I.get("param", "tadv2")))
        R.step(it, mode="full")
        if it in record:
            for n in STEP_FIELDS:
                out[f"{n}_{it}"] = R.get(n)
            for n in LASD_FIELDS:
                out[f"{n}_{it}"] = R.get(n.lower(), module="sgs_param")
    out["dts"] = np.array(dts)
    meta = dict(params=params_record(p), mode="full", record=list(record), cfl=cfl, dyn_init=dyn_init, cs_count=cs_count,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py",
                statements=R.I.nstmt)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: {nsteps} steps, {R.I.nstmt} reference statements")


def run_turbine_case(name="ref_turbines_32x32x8", nsteps=2, eps=0.3, use_rotation=False, tip_speed_ratio=7.0):
    """Rows (f)-3 from the reference text: turbines_forcing (turbines.f90:465-638) through forcing_applied
    (forcing.f90:102-106) and main.f90:263-267 inside two core steps, USE_TURBINES build; two disks (one yawed and
    tilted), adm_correction on.  The node lists / indicator weights are start-up data handed to both sides.
    use_rotation: the ADM with rotation (turbines.f90:607-615), %ind_t and %e_theta handed over as well."""
    from helpers import make_farm
    kw = dict(nx=32, ny=32, Nz=8, lbc_mom=1, ubc_mom=1)
    p = O.Params(**kw)
    R = refrun.Reference(p, files=refrun.TURBINE_FILES, turbines=True)
    farm = make_farm(p)
    for t in farm:
        t.M = 0.9
    if use_rotation:
        assert all(t.ind_t is not None and np.isfinite(t.e_theta).all() for t in farm)
    R.farm_set(farm, eps, adm_correction=True, use_rotation=use_rotation, tip_speed_ratio=tip_speed_ratio)
    u, v, w = initial_fields(p, seed=71)
    u = u + 1.0
    out = {"u0": u, "v0": v, "w0": w}
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    for it in range(1, nsteps + 1):
        R.step(it, mode="core")
    for n in STEP_FIELDS + ("fxa", "fya", "fza"):
        out[f"{n}_{nsteps}"] = R.get(n)
    assert np.count_nonzero(out[f"fxa_{nsteps}"]) > 100 and np.count_nonzero(out[f"fza_{nsteps}"]) > 100
    for n in ("u_d", "u_d_t", "f_n"):
        out["disk_" + n] = np.array(R.farm_get(n))
    for i, t in enumerate(farm):
        out[f"farm{i}_nodes"], out[f"farm{i}_ind"] = np.asarray(t.nodes), np.asarray(t.ind)
        out[f"farm{i}_scalars"] = np.array([t.Ct_prime, t.dia, t.M, t.u_d_T, *t.nhat])
        if use_rotation:
            out[f"farm{i}_ind_t"], out[f"farm{i}_e_theta"] = np.asarray(t.ind_t), np.asarray(t.e_theta)
    meta = dict(params=params_record(p), mode="core", nsteps=nsteps, eps=eps, ndisks=len(farm), adm_correction=True,
                use_rotation=bool(use_rotation), tip_speed_ratio=float(tip_speed_ratio),
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py", statements=R.I.nstmt)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {R.I.nstmt} reference statements")


def run_mpi_lasd_case(name="ref_mpi2_lasd_16x16x8", nproc=2, nsteps=4, seed=53):
    """Rows (f)-2 through the reference's MPI code path: TWO ranks, sgs_model 5 with DYN_init = cs_count = 2, so that
    interpolag_Sdep's halo exchange of F_LM ... F_NN (interpolag_Sdep.f90:244-249), lagrange_Sdep's (:417-420), the
    txz / tyz / tzz halos of sgs_stag and main.f90 and the wall model on rank 0 only all run from the reference text;
    gathered to global fields."""
    kw = dict(nx=16, ny=16, Nz=8, L_x=4.0, L_y=3.0, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, molec=False, dt=2e-3)
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)

    def fn(ref, r):
        for n, g in (("u", ug), ("v", vg), ("w", wg)):
            ref.put(n, O.scatter_slab(g, ref.p))
        for it in range(1, nsteps + 1):
            ref.step(it, mode="full")
        o = {n: ref.get(n) for n in STEP_FIELDS}
        for n in LASD_FIELDS:
            o[n] = ref.get(n.lower(), module="sgs_param")
        o["nstmt"] = ref.I.nstmt
        return o

    t0 = time.time()
    res = refrun.run_ranks(kw, nproc, fn, files=refrun.LASD_FILES, dyn_init=2, cs_count=2)
    ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
    out = {}
    for n in STEP_FIELDS + LASD_FIELDS:
        out[n] = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=n in ("w", "RHSz", "p") + LASD_FIELDS)
    meta = dict(kw=kw, nproc=nproc, nsteps=nsteps, seed=seed, amp=0.3, mode="full", dyn_init=2, cs_count=2,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py, 2 ranks",
                statements=int(sum(res[r]["nstmt"] for r in range(nproc))))
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {meta['statements']} reference statements on {nproc} ranks, {time.time() - t0:.1f} s")


def run_mpi_tavg_case(name="ref_mpi2_tavg_16x16x8", nproc=2, nsteps=2, seed=57):
    """Rows (f)-4 through the reference's MPI code path: tavg%compute (time_average.f90:176-320) on TWO ranks after each
    of two full steps (Smagorinsky, wall model below, stress-free lid), so that the halos inside interp_to_uv_grid /
    interp_to_w_grid (functions.f90:51-141) and the rank-dependent bounds of the accumulators run from the reference
    text; accumulators gathered to global arrays."""
    kw = dict(nx=16, ny=16, Nz=8, L_x=4.0, L_y=3.0, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=1, molec=False)
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)

    def fn(ref, r):
        for n, g in (("u", ug), ("v", vg), ("w", wg)):
            ref.put(n, O.scatter_slab(g, ref.p))
        t = ref.tavg_new()
        for it in range(1, nsteps + 1):
            ref.step(it, mode="full")
            ref.tavg_compute(t, ref.p.dt)
        o = {n: ref.get(n) for n in STEP_FIELDS}
        for n in O.TAVG_FIELDS:
            o["tavg_" + n] = getattr(t, n).a.transpose(2, 1, 0).copy()
        o["total_time"] = float(t.total_time)
        o["nstmt"] = ref.I.nstmt
        return o

    res = refrun.run_ranks(kw, nproc, fn, files=refrun.TAVG_FILES)
    ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
    out = {}
    for n in STEP_FIELDS:
        out[n] = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=n in ("w", "RHSz", "p"))
    for n in O.TAVG_FIELDS:
        out["tavg_" + n] = O.gather_slabs([res[r]["tavg_" + n] for r in range(nproc)], ps, top_extra=False)
    out["total_time"] = np.array(res[0]["total_time"])
    meta = dict(kw=kw, nproc=nproc, nsteps=nsteps, seed=seed, amp=0.3, mode="full",
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py, 2 ranks",
                statements=int(sum(res[r]["nstmt"] for r in range(nproc))))
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {meta['statements']} reference statements on {nproc} ranks")


def run_mpi_turbine_case(name="ref_mpi2_turbines_32x32x8", nproc=2, nsteps=2, eps=0.3, seed=59):
    """Rows (f)-3 through the reference's MPI code path: two disks that span BOTH z slabs of a two-rank run, so that the
    per-rank partial sums and the MPI_Allreduce of the disk velocities (turbines.f90:549-560), the force-field halos
    (:620-622) and interp_to_w_grid across the slab seam run from the reference text, inside two core steps."""
    from helpers import make_farm, farm_for_rank
    kw = dict(nx=32, ny=32, Nz=8, lbc_mom=1, ubc_mom=1)
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)
    ug = ug + 1.0
    farm_g = make_farm(pg)
    for t in farm_g:
        t.M = 0.9
    ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
    assert all(len(farm_for_rank(farm_g, p)[0].ind) > 5 for p in ps), "the first disk must have nodes on both ranks"

    def fn(ref, r):
        ref.farm_set(farm_for_rank(farm_g, ref.p), eps, adm_correction=True)
        for n, g in (("u", ug), ("v", vg), ("w", wg)):
            ref.put(n, O.scatter_slab(g, ref.p))
        for it in range(1, nsteps + 1):
            ref.step(it, mode="core")
        o = {n: ref.get(n) for n in STEP_FIELDS + ("fxa", "fya", "fza")}
        for n in ("u_d", "u_d_t", "f_n"):
            o["disk_" + n] = np.array(ref.farm_get(n))
        o["nstmt"] = ref.I.nstmt
        return o

    res = refrun.run_ranks(kw, nproc, fn, files=refrun.TURBINE_FILES, turbines=True)
    out = {"ug": ug, "vg": vg, "wg": wg}
    for n in STEP_FIELDS + ("fxa", "fya", "fza"):
        out[n] = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=n in ("w", "RHSz", "p", "fza"))
    for n in ("u_d", "u_d_t", "f_n"):
        assert all(np.array_equal(res[r]["disk_" + n], res[0]["disk_" + n]) for r in range(nproc)), n
        out["disk_" + n] = res[0]["disk_" + n]
    for i, t in enumerate(farm_g):
        out[f"farm{i}_nodes"], out[f"farm{i}_ind"] = np.asarray(t.nodes), np.asarray(t.ind)
        out[f"farm{i}_scalars"] = np.array([t.Ct_prime, t.dia, t.M, t.u_d_T, *t.nhat])
    meta = dict(kw=kw, nproc=nproc, nsteps=nsteps, seed=seed, amp=0.3, mode="core", eps=eps, ndisks=len(farm_g),
                adm_correction=True,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py, 2 ranks",
                statements=int(sum(res[r]["nstmt"] for r in range(nproc))))
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {meta['statements']} reference statements on {nproc} ranks")


def run_mpi_rmsdiv_case(name="ref_mpi2_rmsdiv_16x16x8", nproc=2, seed=67):
    """rmsdiv.f90:21-59 as main.f90:367 calls it at the end of a step, on two reference ranks: the L1 divergence metric of
    the derivatives the step formed at :161-172 (i.e. of the synthetic, not divergence-free start field), mpi_reduce to
    rank 0 and the division by nproc."""
    kw = dict(nx=16, ny=16, Nz=8, L_x=4.0, L_y=3.0, lbc_mom=1, ubc_mom=1)
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)
    main = os.path.join(refrun.REF, "main.f90")

    def fn(ref, r):
        ref.I.load(os.path.join(refrun.REF, "rmsdiv.f90"))
        for n, g in (("u", ug), ("v", vg), ("w", wg)):
            ref.put(n, O.scatter_slab(g, ref.p))
        ref.step(1, mode="core")
        # main.f90:367 sits at the end of the step: dudx, dvdy, dwdz are still those of the field the step STARTED from
        v = ref.I.exec_lines(main, 367, 367, ["types", "param", "sim_param"], local={"rmsdivvel": 0.0})
        return {"rms": float(v["rmsdivvel"]), "fields": {n: ref.get(n) for n in ("u", "v", "w")}}

    res = refrun.run_ranks(kw, nproc, fn)
    out = {"ug": ug, "vg": vg, "wg": wg, "rms_rank0": np.array(res[0]["rms"]), "rms_local": np.array([res[r]["rms"] for r in range(nproc)])}
    meta = dict(kw=kw, nproc=nproc, seed=seed,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py, 2 ranks")
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, rms on rank 0 = {res[0]['rms']:.6e}")


def run_filter_kernels(name="ref_filter_kernels_16x32"):
    """test_filter_init (test_filtermodule.f90:38-123) for the three filter types: the kernels G_test (2 Delta) and,
    with sgs_model 5, G_test_test (4 Delta) as the reference builds them, plus one plane filtered with each."""
    from helpers import random_field
    out = {}
    kw = dict(nx=16, ny=32, Nz=4, L_x=4.0, L_y=3.0, sgs=True, sgs_model=5)
    for ifilter in (1, 2, 3):
        p = O.Params(ifilter=ifilter, **kw)
        R = refrun.Reference(p, files=refrun.LASD_FILES)
        I = R.I
        out[f"G_test_{ifilter}"] = np.asarray(I.get("test_filtermodule", "g_test").a).T.copy()
        out[f"G_test_test_{ifilter}"] = np.asarray(I.get("test_filtermodule", "g_test_test").a).T.copy()
        f = random_field(p, 7)
        pl = refrun.F.FArray(np.asfortranarray(f[2].T.copy()), (1, 1))
        I.call("test_filter", pl, module="test_filtermodule")
        out[f"filtered_{ifilter}"] = pl.a.T.copy()
        pl = refrun.F.FArray(np.asfortranarray(f[2].T.copy()), (1, 1))
        I.call("test_test_filter", pl, module="test_filtermodule")
        out[f"filtered2_{ifilter}"] = pl.a.T.copy()
        out["f"] = f
    meta = dict(kw=kw, made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py")
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done")


def run_mpi_case(name="ref_mpi4_full_16x16x8", nproc=4, nsteps=2, seed=51):
    """The reference's MPI code path from the reference text: FOUR ranks (one interpreter per rank, in-process
    mailboxes for mpi_sendrecv / send / recv / allreduce): mpi_sync_real_array halos (mpi_defs.f90:167-264), the
    rank-pipelined tridag_array (tridag_array.f90:22-162), press_stag_array's exchanges and k = 0 chain (:175-246), the tzz
    halo of main.f90:193-197, cfl all-reduce -- two full steps (DNS walls + molecular stress), gathered to global fields."""
    kw = dict(nx=16, ny=16, Nz=8, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, L_x=4.0, L_y=3.0, use_mean_p_force=True,
              mean_p_force_x=1.0, sgs=False, molec=True, nu_molec=1e-2)
    pg = O.Params(nproc=1, **kw)
    ug, vg, wg = O.synthetic_global(pg.nx, pg.ny, pg.Nz, nproc=nproc, seed=seed, amp=0.3, L_x=pg.L_x, L_y=pg.L_y, L_z=pg.L_z)

    def fn(ref, r):
        for n, g in (("u", ug), ("v", vg), ("w", wg)):
            ref.put(n, O.scatter_slab(g, ref.p))
        for it in range(1, nsteps + 1):
            ref.step(it, mode="full")
        o = {n: ref.get(n) for n in STEP_FIELDS}
        o["cfl"] = ref.I.call("get_max_cfl", module="cfl_util")
        o["nstmt"] = ref.I.nstmt
        return o

    res = refrun.run_ranks(kw, nproc, fn)
    ps = [O.Params(nproc=nproc, coord=r, **kw) for r in range(nproc)]
    out = {}
    for n in STEP_FIELDS:
        out[n] = O.gather_slabs([res[r][n] for r in range(nproc)], ps, top_extra=n in ("w", "RHSz", "p"))
    out["max_cfl"] = np.array(res[0]["cfl"])
    assert all(res[r]["cfl"] == res[0]["cfl"] for r in range(nproc))
    meta = dict(kw=kw, nproc=nproc, nsteps=nsteps, seed=seed, amp=0.3, mode="full",
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py, 4 ranks",
                statements=int(sum(res[r]["nstmt"] for r in range(nproc))))
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {meta['statements']} reference statements on {nproc} ranks")


def run_tavg_case(name="ref_tavg_16x16x6", nsteps=2):
    """Rows (f)-4 from the reference text: tavg%compute (time_average.f90:176-320, interp_to_uv_grid / interp_to_w_grid of
    functions.f90) after each of two full DNS steps, with a seeded Cs_opt2 field and growing averaging intervals."""
    from helpers import random_field
    kw = dict(nx=16, ny=16, Nz=6, L_x=4.0, L_y=3.0, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, sgs=False, molec=True, nu_molec=1e-2)
    p = O.Params(**kw)
    R = refrun.Reference(p, files=refrun.TAVG_FILES)
    u, v, w = initial_fields(p, seed=81)
    u = u + 1.0
    cs = 0.01 * np.abs(random_field(p, 82))
    out = {"u0": u, "v0": v, "w0": w, "cs_opt2_0": cs}
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    fa = R.I.get("sgs_param", "cs_opt2")
    fa.a[...] = cs[1:].transpose(2, 1, 0)
    t = R.tavg_new()
    for it in range(1, nsteps + 1):
        R.step(it, mode="full")
        R.tavg_compute(t, p.dt * it)
    for n in O.TAVG_FIELDS:
        out["tavg_" + n] = getattr(t, n).a.transpose(2, 1, 0).copy()
    out["total_time"] = np.array(t.total_time)
    meta = dict(params=params_record(p), mode="full", nsteps=nsteps,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py", statements=R.I.nstmt)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done, {R.I.nstmt} reference statements")


def initial_fields(p, seed=41, amp=0.3):
    u, v, w = O.synthetic_global(p.nx, p.ny, p.Nz, nproc=1, seed=seed, amp=amp, L_x=p.L_x, L_y=p.L_y, L_z=p.L_z)
    return tuple(O.scatter_slab(f, p) for f in (u, v, w))


def params_record(p):
    return {k: getattr(p, k) for k, f in p.__dataclass_fields__.items() if f.init}


def run_step_case(name, kw, mode, record, cfl=None):
    p = O.Params(**kw)
    R = refrun.Reference(p)
    u, v, w = initial_fields(p)
    out = {"u0": u, "v0": v, "w0": w}
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    dts = []
    if cfl is not None:
        # initialize.f90:192-200, from the reference text: dt = get_cfl_dt() * huge (forces an Euler first step)
        R.I.set("param", "cfl", float(cfl)); R.I.set("param", "use_cfl_dt", True); R.I.set("param", "cfl_f", 0.0)
        R.I.exec_lines(os.path.join(refrun.REF, "initialize.f90"), 192, 200, ["types", "param", "sim_param", "cfl_util"],
                       local={"dt_dim": 0.0})
    t0 = time.time()
    for it in range(1, max(record) + 1):
        if cfl is not None:
            R.cfl_dt_step_setup()                                  # main.f90:135-144
            dts.append((R.I.get("param", "dt"), R.I.get("param", "tadv1"), R.I.get("param", "tadv2")))
        R.step(it, mode=mode)
        if it in record:
            for n in STEP_FIELDS:
                out[f"{n}_{it}"] = R.get(n)
    if dts:
        out["dts"] = np.array(dts)
    meta = dict(params=params_record(p), mode=mode, record=list(record), cfl=cfl,
                made_by="oracle/make_reference_fixtures.py: reference sources interpreted by oracle/f90exec.py",
                statements=R.I.nstmt)
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: {max(record)} steps, {R.I.nstmt} reference statements, {time.time() - t0:.1f} s")


def run_routines(name="ref_routines_16x16x6"):
    """Routine by routine on seeded random fields: ddx, ddy, ddxy, filt_da, ddz_uv, ddz_w, padd / unpadd, convec for the
    four wall-condition combinations the tests use, press_stag_array, test_filter, get_max_cfl / get_cfl_dt."""
    from helpers import random_field
    out = {}
    kw = dict(nx=16, ny=16, Nz=6, L_x=4.0, L_y=3.0)
    p = O.Params(**kw)
    R = refrun.Reference(p)
    I = R.I
    f = random_field(p, 1)
    out["f"] = f
    D = "derivatives"
    uses = ["types", "param", "sim_param", "derivatives", "fft", "test_filtermodule", "cfl_util"]
    main_like = lambda text: None
    # the calls are made exactly as main.f90 / divstress make them: module arrays of sim_param as actual arguments
    R.put("u", f)
    I.call("ddx", I.get("sim_param", "u"), I.get("sim_param", "dudx"), 0, module=D); out["ddx"] = R.get("dudx")
    I.call("ddy", I.get("sim_param", "u"), I.get("sim_param", "dudy"), 0, module=D); out["ddy"] = R.get("dudy")
    I.call("ddxy", I.get("sim_param", "u"), I.get("sim_param", "dvdx"), I.get("sim_param", "dvdy"), 0, module=D)
    out["ddxy_x"], out["ddxy_y"] = R.get("dvdx"), R.get("dvdy")
    I.call("ddz_uv", I.get("sim_param", "u"), I.get("sim_param", "dudz"), 0, module=D); out["ddz_uv"] = R.get("dudz")
    I.call("ddz_w", I.get("sim_param", "u"), I.get("sim_param", "dwdz"), 0, module=D); out["ddz_w"] = R.get("dwdz")
    I.call("filt_da", I.get("sim_param", "u"), I.get("sim_param", "dudx"), I.get("sim_param", "dudy"), 0, module=D)
    out["filt_da_f"], out["filt_da_x"], out["filt_da_y"] = R.get("u"), R.get("dudx"), R.get("dudy")
    # test_filter(f) on one plane, in place (test_filtermodule.f90:126-146)
    pl = refrun.F.FArray(np.asfortranarray(f[2].T.copy()), (1, 1))
    I.call("test_filter", pl, module="test_filtermodule")
    out["test_filter_plane2"] = pl.a.T.copy()
    # cfl_util on a seeded state
    u, v, w = initial_fields(p, seed=91)
    for n, a in (("u", u), ("v", v), ("w", w)):
        R.put(n, a)
    out["cfl_u"], out["cfl_v"], out["cfl_w"] = u, v, w
    I.set("param", "cfl", 0.0625)
    out["max_cfl"] = np.array(I.call("get_max_cfl", module="cfl_util"))
    out["cfl_dt"] = np.array(I.call("get_cfl_dt", module="cfl_util"))
    # convec and press_stag_array for several wall conditions
    for tag, bc in (("11d", (1, 1, False)), ("00d", (0, 0, False)), ("22l", (2, 2, True)), ("10l", (1, 0, True))):
        pc = O.Params(lbc_mom=bc[0], ubc_mom=bc[1], sgs=bc[2], **kw)
        Rc = refrun.Reference(pc, files=[x for x in refrun.FILES if x not in ("sgs_stag_util.f90", "wallstress.f90", "divstress_uv.f90", "divstress_w.f90")])
        names = ("u", "v", "w", "dudy", "dudz", "dvdx", "dvdz", "dwdx", "dwdy")
        for i, n in enumerate(names):
            Rc.put(n, random_field(pc, 20 + i))
        Rc.call("convec")
        for n in ("RHSx", "RHSy", "RHSz"):
            out[f"convec_{tag}_{n}"] = Rc.get(n)
    pp = O.Params(**kw)
    Rp = refrun.Reference(pp)
    u, v, w = initial_fields(pp, seed=31)
    for n, a in (("u", u), ("v", v), ("w", w)):
        Rp.put(n, a)
    dz = 0.1 * random_field(pp, 32)
    Rp.put("divtz", dz)
    out["press_u"], out["press_v"], out["press_w"], out["press_divtz"] = u, v, w, dz
    Rp.call("press_stag_array")
    for n in ("p", "dpdx", "dpdy", "dpdz"):
        out[f"press_{n}"] = Rp.get(n)
    meta = dict(params=params_record(p), made_by="oracle/make_reference_fixtures.py")
    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: done")


if __name__ == "__main__":
    if not refrun.available():
        sys.exit("the reference sources are not at " + refrun.REF)
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    if not only or "routines" in only:
        run_routines()
    for name, c in STEP_CASES.items():
        if not only or name in only:
            run_step_case(name, **c)
    if not only or "lasd" in only:
        run_lasd_case()
    if not only or "lasd_cfl" in only:
        run_lasd_cfl_case()
    if not only or "tavg" in only:
        run_tavg_case()
    if not only or "turbines" in only:
        run_turbine_case()
    if not only or "turbines_rot" in only:
        run_turbine_case(name="ref_turbines_rot_32x32x8", use_rotation=True, tip_speed_ratio=5.5)
    if not only or "mpi" in only:
        run_mpi_case()
    if not only or "filters" in only:
        run_filter_kernels()
    if not only or "mpi_lasd" in only:
        run_mpi_lasd_case()
    if not only or "mpi_tavg" in only:
        run_mpi_tavg_case()
    if not only or "mpi_turbines" in only:
        run_mpi_turbine_case()
    if not only or "mpi_rmsdiv" in only:
        run_mpi_rmsdiv_case()
