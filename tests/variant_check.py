"""Runs the routine- and step-level parity checks in THIS process (whose environment selects a
kernel variant: LESGO_XW / LESGO_PROD_CHUNK / LESGO_REUSE are read once per process).
Used by test_emul_parity.py (emulator, CPU) and test_gpu_variants.py (sm_100a build, B200)."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import lesgo_b200  # noqa: E402
from helpers import (O, check_convec, check_derivatives, check_lasd_steps, check_press, check_steps,  # noqa: E402
                     emul_library, make_dims)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emul", action="store_true")
    ap.add_argument("--what", default="deriv,convec,press,steps,full")
    ap.add_argument("--grid", default="16,16,6")
    a = ap.parse_args()
    nx, ny, Nz = (int(x) for x in a.grid.split(","))
    what = a.what.split(",")

    def core(p):
        if a.emul:
            return lesgo_b200.Core(make_dims(p), lib=emul_library())
        return lesgo_b200.Core(make_dims(p, device=0))

    worst = 0.0
    if "deriv" in what:
        p = O.Params(nx=nx, ny=ny, Nz=Nz, L_x=4.0, L_y=3.0)
        worst = max(worst, max(check_derivatives(core(p), p).values()))
    if "convec" in what:
        for bc in [(1, 1, False), (0, 0, False), (2, 2, True), (1, 0, True)]:
            p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=bc[0], ubc_mom=bc[1], sgs=bc[2])
            worst = max(worst, max(check_convec(core(p), p).values()))
    if "press" in what:
        p = O.Params(nx=nx, ny=ny, Nz=Nz)
        worst = max(worst, max(check_press(core(p), p).values()))
    if "steps" in what:
        p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=1, ubc_mom=1, utop=0.5, ubot=-0.5, use_mean_p_force=True, mean_p_force_x=1.0)
        worst = max(worst, max(check_steps(core(p), p, nsteps=2, tol=1e-11).values()))
    if "full" in what:
        p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=2, ubc_mom=2, sgs=True, sgs_model=1, molec=False, use_mean_p_force=True,
                     mean_p_force_x=1.0)
        worst = max(worst, max(check_steps(core(p), p, nsteps=2, tol=1e-11, mode="full").values()))
    if "lasd" in what:
        p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=2, ubc_mom=0, sgs=True, sgs_model=5, dt=2e-3)
        worst = max(worst, max(check_lasd_steps(core(p), p, nsteps=4, tol=1e-11).values()))
    print(f"variant_check ok worst={worst:.3e}")


if __name__ == "__main__":
    main()
