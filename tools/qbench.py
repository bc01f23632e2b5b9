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
ap.add_argument("--lasd", type=int, default=0, help="cs_count: full step with sgs_model 5, lagrange_Sdep every cs_count steps")
a = ap.parse_args()
nx, ny, Nz = (int(x) for x in a.grid.split(","))
dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, device=0, sgs=bool(a.lasd), lbc_mom=2 if a.lasd else 1, ubc_mom=0 if a.lasd else 1)
core = lesgo_b200.Core(dims)
u, v, w = synthetic_slab(dims)
for n, arr in (("u", u), ("v", v), ("w", w)):
    core.upload(n, arr)
for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
    core.upload(n, np.zeros(dims.shape))
kw = dict(dt=2e-4, tadv1=1.5, tadv2=-0.5, mode=a.mode, ubot=-1.0, utop=1.0)
if a.lasd:
    # LES half channel with the Lagrangian scale-dependent model: one lagrange_Sdep per timed step
    # (the amortised cost per step is that share / cs_count)
    for n in ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
        core.upload(n, np.zeros(dims.shape))
    kw.update(mode=1, sgs_model=5, lagran_dt=a.lasd * 2e-4)
    core.step(first_step=True, lasd_cs_init=True, **kw)
    core.step(lasd_update=True, lasd_init_F=True, **kw)
    kw.update(lasd_update=True)
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
