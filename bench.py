#!/usr/bin/env python
"""bench.py -- grid-point updates per second of the LESGO per-timestep core on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # CPU reference arm (oracle port)

One "step" = one pass of the hot path: filt_da x3, ddz_uv x2, ddz_w, wall
derivatives, convec, RHS assembly, AB2, press_stag_array (+ tridag), RHS -= grad p,
project -- main.f90:155-344 with the stress divergence taken as zero (the SURVEY 8(d) core;
the complete step with wall stress, Smagorinsky stress and its divergence is timed
separately as `full_step`) -- on device-resident synthetic channel fields.  Metric: Mpts/s with
points = nx*ny*(nz_tot-1), whole job.  The workload is the 512x512x256 channel at every
N (strong scaling along LESGO's own z-slab decomposition); it fits one B200.

JSON keys follow the driver contract; `roofline` is the whole-step figure of SURVEY
8(d): achieved = 296 B/point/step * points / step time, against the measured HBM copy
bandwidth in MEASURED_PEAKS.json; `kernels` adds a per-pass breakdown (CUDA events
around every launch, separate instrumented step, not inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_CORE_BYTES = 296.0          # 37 FP64 words per point per step
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", default="512,512,256", help="nx,ny,Nz (lesgo.conf Nz)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the LASD and actuator-disk timings (rows (f)-2, (f)-3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=1)
    return ap.parse_args()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class _Disk:
    pass


def synthetic_farm(dims, rows=4, cols=6):
    """rows x cols actuator disks (diameter 0.1 L_y... as in test-cases/turbines_ADM: hub height 0.1 L_z-ish,
    unit normal -x) with smooth indicator weights normalised to unit volume integral: the node lists
    turbines_nodes (turbines.f90:275-462) would hand to lesgo_gpu_turbines_init, for this rank's slab."""
    nx, ny, nz = dims.nx, dims.ny, dims.nz
    dx, dy, dz = dims.L_x / nx, dims.L_y / ny, dims.dz
    dia = 0.1 * dims.L_y
    height = 0.25 * dims.L_z
    thk = max(1.5 * dx, 0.1 * dia)
    base = dims.coord * (nz - 1)
    farm = []
    for r in range(rows):
        for c_ in range(cols):
            xl, yl = (c_ + 0.5) * dims.L_x / cols, (r + 0.5) * dims.L_y / rows
            ic, jc, kc = int(round(xl / dx)), int(round(yl / dy)), int(round(height / dz + 0.5))
            hi_, hj, hk = int(thk / dx) + 2, int(0.6 * dia / dy) + 2, int(0.6 * dia / dz) + 2
            ii, jj, kk = np.meshgrid(np.arange(ic - hi_, ic + hi_ + 1), np.arange(jc - hj, jc + hj + 1),
                                     np.arange(max(kc - hk, 1), min(kc + hk, dims.nz_tot - 1) + 1), indexing="ij")
            rx, ry, rz = (ii - 1) * dx - xl, (jj - 1) * dy - yl, (kk - 0.5) * dz - height
            wgt = np.exp(-(np.sqrt(ry ** 2 + rz ** 2) / (0.5 * dia)) ** 8) * np.exp(-(rx / (0.5 * thk)) ** 4)
            keep = wgt > 1e-2
            wsum = float(wgt[keep].sum() * dx * dy * dz)
            mine = keep & (kk >= base + 1) & (kk <= base + nz - 1)
            t = _Disk()
            t.nodes = np.stack([(ii[mine] - 1) % nx + 1, (jj[mine] - 1) % ny + 1, kk[mine] - base], axis=1).astype(np.int32)
            t.ind = (wgt[mine] / wsum).astype(np.float64)
            t.nhat, t.Ct_prime, t.dia, t.M, t.u_d_T = (-1.0, 0.0, 0.0), 1.33, dia, 0.9, -1.0
            farm.append(t)
    return farm


def synthetic_slab(dims, seed=20240607):
    """Channel-like fields for this rank: parabolic mean + tapered uniform noise, w = 0 on
    the walls, ghost planes consistent with the neighbours (global planes generated
    deterministically per level so every rank sees the same values)."""
    nx, ny, nz, ld = dims.nx, dims.ny, dims.nz, dims.ld
    base = dims.coord * (nz - 1)
    out = [np.zeros(dims.shape) for _ in range(3)]
    for k in range(nz + 1):
        g = base + k                     # global level
        if g < 1 or g > dims.nz_tot:
            for a in out:
                a[k] = -1234567890.0
            continue
        rng = np.random.default_rng([seed, g])
        zuv = (g - 0.5) * dims.dz
        zw = (g - 1.0) * dims.dz
        for comp, a in enumerate(out):
            noise = rng.random((ny, nx)) - 0.5
            z = zw if comp == 2 else zuv
            taper = math.sqrt(max(math.sin(math.pi * min(max(z / dims.L_z, 0.0), 1.0)), 0.0))
            a[k, :, :nx] = 0.3 * noise * (taper if comp else 1.0)
            if comp == 0:
                a[k, :, :nx] += 1.5 * (1.0 - (zuv / (0.5 * dims.L_z) - 1.0) ** 2)
        if g == 1 or g == dims.nz_tot:
            out[2][k] = 0.0
    return out


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML every ~5 ms when
    pynvml is importable, else `nvidia-smi` polling (the recipe's query line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self._stop_evt = index, threading.Event()
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            self.h = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if index < len(ids) else index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                            "-i", str(self.index)], capture_output=True, text=True, timeout=5)
        if r.returncode == 0 and r.stdout.strip():
            row = [x.strip() for x in r.stdout.strip().split(",")]
            self.sm.append(float(row[1])); self.mx.append(float(row[2]))
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], row[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.005 if self.nvml else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "power_w_max": max(self.power) if self.power else None,
                "how": "pynvml, 5 ms" if self.nvml else "nvidia-smi polling"}


def cpu_core_step_rate(nx, ny, workers, budget_s=20.0):
    """Time the oracle's core step (the CPU restatement of the reference path) on a
    bounded z-sample of the same (nx, ny) grid; returns (Mpts/s, sample text, planes)."""
    from oracle import lesgo_oracle as O
    O.FFT_WORKERS = workers
    Nz = 8
    p = O.Params(nx=nx, ny=ny, Nz=Nz, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0)
    sp = O.Spectral(p)
    s = O.State(p)
    rng = np.random.default_rng(1)
    for n in ("u", "v", "w"):
        getattr(s, n)[:, :, :nx] = rng.standard_normal((p.nz + 1, ny, nx))
    comm = O.LocalComm()
    O.step(s, sp, comm, mode="core", first_step=True)       # warm-up (pocketfft plan cache)
    t0 = time.perf_counter()
    n = 0
    while True:
        O.step(s, sp, comm, mode="core")
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    dt = (time.perf_counter() - t0) / n
    pts = nx * ny * (p.nz_tot - 1)
    return pts / dt / 1e6, f"{nx}x{ny}x{Nz} z-sample of the workload, {n} core steps, scipy.fft workers={workers}", dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    nx, ny, Nz = (int(x) for x in args.grid.split(","))
    cores = os.cpu_count() or 1
    # K steps, each a bounded sample
    v, sample, dt = cpu_core_step_rate(nx, ny, cores, budget_s=max(5.0, min(60.0, 4.0 * (args.steps + args.warmup))))
    line = {"impl": "reference", "metric": "grid-point updates/sec", "value": v, "unit": "Mpts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"LES channel core step {nx}x{ny}x{Nz} (CPU: bounded z-sample)", "grid": [nx, ny, Nz]},
            "cpu_baseline": {"value": v, "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference cannot be built here (no gfortran/FFTW3/MPI): this is the oracle port of its "
                    "algorithm with multi-threaded pocketfft, per-step time %.3f s on the sample" % dt}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import lesgo_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, Nz = (int(x) for x in args.grid.split(","))
    dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, nproc=world, coord=rank, lbc_mom=1, ubc_mom=1, sgs=True, device=local)
    core = lesgo_b200.Core(dims)
    stream = torch.cuda.current_stream()
    core.set_stream(stream.cuda_stream)
    if world > 1:
        from lesgo_b200 import slab
        slab.bootstrap_comm(core, dist)
    dt, tadv1, tadv2 = 2e-4, 1.5, -0.5
    u, v, w = synthetic_slab(dims)
    for n, a in (("u", u), ("v", v), ("w", w)):
        core.upload(n, a)
    zero = np.zeros(dims.shape)
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, zero)
    step_kw = dict(dt=dt, tadv1=tadv1, tadv2=tadv2, mode=0, ubot=-1.0, utop=1.0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    core.step(first_step=True, **step_kw)
    for _ in range(max(args.warmup - 1, 2)):
        core.step(**step_kw)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = core.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        core.step(**step_kw)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = core.launch_count - l0
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cfl = core.max_cfl(dt)
    if not math.isfinite(cfl) or cfl <= 0.0 or cfl > 10.0:
        raise SystemExit(f"bench.py: simulation state is not sane after the timed steps (CFL = {cfl})")
    ms_step = ms / args.steps
    points = nx * ny * (dims.nz_tot - 1)
    value = points / (ms_step * 1e-3) / 1e6

    # per-pass breakdown: one instrumented step outside the timed region
    core.profile(True)
    core.step(**step_kw)
    kern = core.profile(False, report=True)
    peak, peak_src = hbm_peak()
    achieved = A_CORE_BYTES * points / (ms_step * 1e-3) / 1e9 / world
    traffic, traffic_src = None, None
    try:
        tpath = os.path.join(ROOT, "profiles", "r3_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("grid") == [nx, ny, Nz] and world == 1:
            traffic, traffic_src = tj["dram_bytes_per_step"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": A_CORE_BYTES * points,
                "peak_source": peak_src,
                "definition": "296 B/point/step (SURVEY 8d A_core) * points / step time / n_gpus; whole hot path"}
    kernels = {k: {"launches": n, "ms": round(t, 4)} for k, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1])}

    # the complete step (rows (f)-1 on the device too: equilibrium wall model, Smagorinsky stress,
    # stress divergence), timed separately; the headline stays the core step of SURVEY 8(d)
    full = None
    try:
        fkw = dict(step_kw, mode=1, sgs_model=1, nu=1e-4)
        for _ in range(2):
            core.step(**fkw)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        nfull = max(2, args.steps // 2)
        for _ in range(nfull):
            core.step(**fkw)
        f1.record(stream)
        barrier()
        fms = f0.elapsed_time(f1) / nfull
        if dist is not None:
            t = torch.tensor([fms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fms = float(t.item())
        full = {"ms_per_step": fms, "value": points / (fms * 1e-3) / 1e6, "unit": "Mpts/s",
                "what": "core + DNS-wall wallstress + calc_Sij + Smagorinsky sgs_stag + divstress_uv/w"}
    except Exception as e:  # noqa
        full = {"error": str(e)}

    # rows (f)-2 and (f)-3 of SURVEY section 8, timed the same way on the same grid: the Lagrangian
    # scale-dependent model (one lagrange_Sdep per timed step; the reference runs it every cs_count = 5
    # steps) and a 4 x 6 array of actuator disks
    def timed_steps(kw, n):
        for _ in range(2):
            core.step(**kw)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            core.step(**kw)
        b.record(stream)
        barrier()
        t_ms = a.elapsed_time(b) / n
        if dist is not None:
            tt = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ms = float(tt.item())
        return t_ms

    lasd = None
    if not args.no_extras:
        try:
            lkw = dict(step_kw, mode=1, sgs_model=5, nu=1e-4, lagran_dt=5 * dt)
            core.step(lasd_cs_init=True, **lkw)
            core.step(lasd_update=True, lasd_init_F=True, **lkw)
            base_ms = timed_steps(lkw, 3)
            upd_ms = timed_steps(dict(lkw, lasd_update=True), 3)
            lasd = {"ms_per_step_with_update": upd_ms, "ms_per_step_without": base_ms,
                    "ms_per_step_cs_count_5": base_ms + (upd_ms - base_ms) / 5.0,
                    "value_cs_count_5": points / ((base_ms + (upd_ms - base_ms) / 5.0) * 1e-3) / 1e6, "unit": "Mpts/s",
                    "what": "full step with sgs_model 5: + interpolag_Sdep + 42 test filters per plane + running averages"}
        except Exception as e:  # noqa
            lasd = {"error": str(e)}
    turb = None
    if not args.no_extras:
        try:
            farm = synthetic_farm(dims)
            core.turbines_init(farm)
            tkw = dict(step_kw, mode=1, sgs_model=1, nu=1e-4)
            t_ms = timed_steps(dict(tkw, turbines=True, turbines_eps=0.1), 3)
            turb = {"ms_per_step": t_ms, "value": points / (t_ms * 1e-3) / 1e6, "unit": "Mpts/s", "disks": len(farm),
                    "nodes_this_rank": int(sum(len(t.ind) for t in farm)),
                    "what": "full step (Smagorinsky) + turbines_forcing: gather, all-reduce, scatter, RHS += f"}
        except Exception as e:  # noqa
            turb = {"error": str(e)}
    tavg = None
    if not args.no_extras:
        try:
            core.tavg_compute(dt)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3):
                core.tavg_compute(dt)
            b.record(stream)
            barrier()
            t_ms = a.elapsed_time(b) / 3
            # 17 fields + 5 interpolated read, 26 accumulators read and written, 5 interpolated written, per point
            tavg = {"ms_per_call": t_ms, "GBps": (17 + 5 + 2 * 26 + 5 + 7) * 8 * points / world / (t_ms * 1e-3) / 1e9,
                    "what": "tavg%compute: interpolations + 26 accumulators (row (f)-4), algorithmic bytes / time"}
        except Exception as e:  # noqa
            tavg = {"error": str(e)}

    e2e = None
    if rank == 0 and not args.no_e2e and world == 1:
        e2e = run_e2e(core, dims, u, v, w, dt, tadv1, args.e2e_steps, points)
    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        cv, sample, _ = cpu_core_step_rate(nx, ny, cores, budget_s=15.0)
        cpu = {"value": cv, "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        line = {"metric": "grid-point updates/sec", "value": value, "unit": "Mpts/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"LES channel core timestep {nx}x{ny}x{Nz} FP64 (device-resident)",
                           "grid": [nx, ny, Nz], "decomposition": f"z-slabs x{world}",
                           "pressure_transposes": ("n/a" if world == 1 else
                                                   ("NVLink peer-memory stores" if getattr(core, "p2p_enabled", False)
                                                    else "NCCL all-to-all")),
                           "l2": "inputs larger than L2 (%.0f MB per field)" % (np.prod(dims.shape) * 8 / 1e6)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks, "kernels": kernels, "max_cfl": cfl, "full_step": full, "lasd_step": lasd, "turbines_step": turb, "tavg": tavg}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(core, dims, u, v, w, dt, tadv1, nsteps, points):
    """The same hot path through the reference-facing per-routine C ABI with HOST buffers
    (pinned), i.e. what the Fortran shim does when LESGO keeps its module arrays on the
    host: every call stages its inputs H2D and its outputs D2H inside the timed region."""
    import torch

    def pinned(a=None):
        t = torch.empty(dims.shape, dtype=torch.float64, pin_memory=True)
        if a is None:
            t.zero_()
        else:
            t.copy_(torch.from_numpy(a))
        return t.numpy()

    F = {n: pinned() for n in ("dudx", "dudy", "dudz", "dvdx", "dvdy", "dvdz", "dwdx", "dwdy", "dwdz", "RHSx", "RHSy",
                               "RHSz", "divtz", "p", "dpdx", "dpdy", "dpdz")}
    F["u"], F["v"], F["w"] = pinned(u), pinned(v), pinned(w)
    nb = float(np.prod(dims.shape) * 8)
    h2d = nb * (3 * 1 + 3 * 1 + 9 + 4)          # filt_da in; ddz in; convec in; press in (outputs are not uploaded)
    d2h = nb * (3 * 3 + 3 + 3 + 4)

    def one():
        core.filt_da(F["u"], F["dudx"], F["dudy"])
        core.filt_da(F["v"], F["dvdx"], F["dvdy"])
        core.filt_da(F["w"], F["dwdx"], F["dwdy"])
        core.ddz_uv(F["u"], F["dudz"])
        core.ddz_uv(F["v"], F["dvdz"])
        core.ddz_w(F["w"], F["dwdz"])
        core.convec(F["u"], F["v"], F["w"], F["dudy"], F["dudz"], F["dvdx"], F["dvdz"], F["dwdx"], F["dwdy"],
                    F["RHSx"], F["RHSy"], F["RHSz"])
        core.press_stag_array(F["u"], F["v"], F["w"], F["divtz"], dt, tadv1, F["p"], F["dpdx"], F["dpdy"], F["dpdz"])

    one()                                   # warm-up (allocates the staging buffers)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        one()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / nsteps
    out = {"value": points / t / 1e6, "unit": "Mpts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": t * 1e3, "api": "per-routine C ABI (filt_da x3, ddz_uv x2, ddz_w, convec, press_stag_array), "
           "pinned host arrays"}
    # for comparison, NOT the headline: the whole-step entry with the state held by the host -- upload
    # u, v, w, RHSx, RHSy, RHSz, one lesgo_gpu_step, download u, v, w, p, RHSx, RHSy, RHSz, every step
    try:
        S = {n: pinned() for n in ("RHSx", "RHSy", "RHSz", "p")}
        S["u"], S["v"], S["w"] = pinned(u), pinned(v), pinned(w)
        kw = dict(dt=dt, tadv1=tadv1, tadv2=-0.5, mode=0, ubot=-1.0, utop=1.0)

        def one_step(first=False):
            for n in ("u", "v", "w", "RHSx", "RHSy", "RHSz"):
                core.upload(n, S[n])
            core.step(first_step=first, **kw)
            for n in ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz"):
                core.download(n, S[n])

        one_step(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            one_step()
        torch.cuda.synchronize()
        ts = (time.perf_counter() - t0) / nsteps
        out["whole_step_api"] = {"value": points / ts / 1e6, "unit": "Mpts/s", "ms_per_step": ts * 1e3,
                                 "h2d_bytes_per_step": 6 * nb, "d2h_bytes_per_step": 7 * nb,
                                 "api": "lesgo_gpu_upload x6 + lesgo_gpu_step + lesgo_gpu_download x7, pinned host arrays"}
    except Exception as e:  # noqa
        out["whole_step_api"] = {"error": str(e)}
    return out


if __name__ == "__main__":
    main()
