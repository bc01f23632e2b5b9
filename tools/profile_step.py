#!/usr/bin/env python
"""Run one device-resident core step between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly the hot-path kernels of one step."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lesgo_b200  # noqa: E402
from bench import synthetic_slab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", default="512,512,16")
ap.add_argument("--what", default="step", choices=["step", "filt_da", "convec", "press"])
ap.add_argument("--lasd", action="store_true", help="profile a full step (mode 1, sgs_model 5) that runs lagrange_Sdep")
a = ap.parse_args()
nx, ny, Nz = (int(x) for x in a.grid.split(","))
dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, device=0, sgs=a.lasd, lbc_mom=2 if a.lasd else 1, ubc_mom=0 if a.lasd else 1)
core = lesgo_b200.Core(dims)
u, v, w = synthetic_slab(dims)
for n, arr in (("u", u), ("v", v), ("w", w)):
    core.upload(n, arr)
for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
    core.upload(n, np.zeros(dims.shape))
kw = dict(dt=2e-4, tadv1=1.5, tadv2=-0.5, mode=0, ubot=-1.0, utop=1.0)
if a.lasd:
    for n in ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
        core.upload(n, np.zeros(dims.shape))
    kw.update(mode=1, sgs_model=5, lagran_dt=5 * 2e-4)
    core.step(first_step=True, lasd_cs_init=True, **kw)
    core.step(lasd_update=True, lasd_init_F=True, **kw)
    kw.update(lasd_update=True)
core.step(first_step=True, **kw)
core.step(**kw)
core.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
core.step(**kw)
core.synchronize()
rt.cudaProfilerStop()
print("profiled one step; launches so far:", core.launch_count)
