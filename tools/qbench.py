#!/usr/bin/env python
"""Quick device-resident timing of the core step with the per-pass breakdown (developer tool)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lesgo_b200  # noqa: E402
from bench import synthetic_slab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", default="512,512,256")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
nx, ny, Nz = (int(x) for x in a.grid.split(","))
dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, device=0)
core = lesgo_b200.Core(dims)
u, v, w = synthetic_slab(dims)
for n, arr in (("u", u), ("v", v), ("w", w)):
    core.upload(n, arr)
for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
    core.upload(n, np.zeros(dims.shape))
kw = dict(dt=2e-4, tadv1=1.5, tadv2=-0.5, mode=a.mode, ubot=-1.0, utop=1.0)
core.step(first_step=True, **kw)
for _ in range(2):
    core.step(**kw)
core.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    core.step(**kw)
core.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / a.steps
core.profile(True)
core.step(**kw)
kern = core.profile(False, report=True)
tot = sum(t for _, t in kern.values())
print(f"grid {a.grid} mode {a.mode} XW={os.environ.get('LESGO_XW', '1')}: {ms:.3f} ms/step (profiled sum {tot:.3f})")
for k, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:14s} {n:3d} launches {t:8.4f} ms")
