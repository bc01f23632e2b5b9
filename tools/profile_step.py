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
a = ap.parse_args()
nx, ny, Nz = (int(x) for x in a.grid.split(","))
dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, device=0)
core = lesgo_b200.Core(dims)
u, v, w = synthetic_slab(dims)
for n, arr in (("u", u), ("v", v), ("w", w)):
    core.upload(n, arr)
for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
    core.upload(n, np.zeros(dims.shape))
kw = dict(dt=2e-4, tadv1=1.5, tadv2=-0.5, mode=0, ubot=-1.0, utop=1.0)
core.step(first_step=True, **kw)
core.step(**kw)
core.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
core.step(**kw)
core.synchronize()
rt.cudaProfilerStop()
print("profiled one step; launches so far:", core.launch_count)
