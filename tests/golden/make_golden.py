#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

These fixtures are produced by the ORACLE (oracle/lesgo_oracle.py), not by the reference: the
reference (Fortran + FFTW3 + MPI) cannot be built or run in the build container and ships no
golden vectors for this path ("parity unpinned").  They pin the oracle against
drift (tests/test_golden.py, CPU) and give the CUDA path a second, frozen target (-m gpu).
Inputs are regenerated from the seed; outputs after `nsteps` full steps are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import lesgo_oracle as O  # noqa: E402

CASES = {
    "dns_couette_32x32x8": dict(kw=dict(nx=32, ny=32, Nz=8, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0, sgs=False,
                                        molec=True, nu_molec=1e-3, L_x=4 * np.pi), seed=101, nsteps=2, mode="full"),
    "les_channel_32x16x8": dict(kw=dict(nx=32, ny=16, Nz=8, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False,
                                        use_mean_p_force=True, mean_p_force_x=1.0), seed=102, nsteps=2, mode="full"),
    "core_halfchannel_16x16x6": dict(kw=dict(nx=16, ny=16, Nz=6, lbc_mom=1, ubc_mom=0), seed=103, nsteps=3, mode="core"),
    # rows (f)-2..4: Lagrangian scale-dependent model (DYN_init = cs_count = 2) + two actuator disks + time averages
    "lasd_turbines_halfchannel_32x16x12": dict(kw=dict(nx=32, ny=16, Nz=12, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5,
                                                       dt=4e-3), seed=104, nsteps=4, mode="full", lasd=True, turbines=True),
}
NAMES = ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz")
EXTRA = ("Cs_opt2", "F_LM", "F_QN")            # model state (planes 1..nz) of the lasd cases
TAVG = ("uw", "p", "fx", "cs_opt2")             # accumulators (planes 1..nz-1) of the lasd cases


def run_case(case):
    p = O.Params(**case["kw"])
    sp = O.Spectral(p)
    G = O.test_filter_kernel(sp)
    u, v, w = O.synthetic_global(p.nx, p.ny, p.Nz, seed=case["seed"], amp=0.3, L_x=p.L_x, L_y=p.L_y, L_z=p.L_z)
    s = O.State(p)
    s.u, s.v, s.w = (O.scatter_slab(f, p) for f in (u, v, w))
    farm = tavg = G2 = None
    if case.get("lasd"):
        from helpers import lasd_schedule, make_farm
        G2 = O.test_filter_kernel(sp, alpha=4.0)
        farm = make_farm(p)
        tavg = O.Tavg(p)
    for it in range(case["nsteps"]):
        lasd = None
        if case.get("lasd"):
            sch = lasd_schedule(p, it)
            lasd = dict(sp=sp, G_test=G, G_test_test=G2, lagran_dt=sch["lagran_dt"], cs_init=sch["lasd_cs_init"],
                        update=sch["lasd_update"], init_F=sch["lasd_init_F"])
        O.step(s, sp, O.LocalComm(), mode=case["mode"], first_step=(it == 0), G_test=G, lasd=lasd,
               turbines=dict(farm=farm, eps=0.3) if farm else None)
        if tavg is not None:
            O.tavg_compute(tavg, s, p, O.LocalComm(), p.dt, forces=True)
    out = {n: getattr(s, n)[1:p.nz + (1 if n in ("w", "p", "RHSz") else 0), :, :p.nx].copy() for n in NAMES}
    if case.get("lasd"):
        out.update({n: getattr(s, n)[1:p.nz + 1, :, :p.nx].copy() for n in EXTRA})
        out.update({"tavg_" + n: getattr(tavg, n)[1:p.nz].copy() for n in TAVG})
        out["u_d_T"] = np.array([t.u_d_T for t in farm])
    return p, out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, case in CASES.items():
        p, out = run_case(case)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
