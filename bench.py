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
A_STEP_BYTES = 976.0          # the whole device-resident step in the reference's dataflow (122 words)
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback
# BASELINE.json configs[3] (headline), configs[2], configs[4]
WORKLOADS = {
    "core512": dict(grid=(512, 512, 256), kind="core", text="LES channel core timestep 512x512x256 FP64 (device-resident)"),
    "lasd256": dict(grid=(256, 256, 128), kind="lasd", text="LES channel 256x256x128, full step with the Lagrangian "
                    "scale-dependent SGS model (lagrange_Sdep every cs_count = 5 steps), FP64, device-resident"),
    "adm1024": dict(grid=(1024, 512, 256), kind="adm", text="turbines_ADM wind-farm channel 1024x512x256, full LES step "
                    "(Smagorinsky) with 24 actuator disks, FP64, device-resident"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="core512", choices=sorted(WORKLOADS),
                    help="core512 (default, the headline of BASELINE.json): core step at 512x512x256; lasd256: configs[2], "
                         "256x256x128 full LES step with the Lagrangian scale-dependent model (update every cs_count = 5 "
                         "steps); adm1024: configs[4], 1024x512x256 full LES step with 24 actuator disks")
    ap.add_argument("--grid", default=None, help="nx,ny,Nz (lesgo.conf Nz); overrides the workload's grid")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank-vs-oracle parity check before the timing")
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


def grid_of(args):
    if args.grid:
        return tuple(int(x) for x in args.grid.split(","))
    return WORKLOADS[args.workload]["grid"]


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm for this path on the box's host cores.  The reference itself (Fortran +
    FFTW3 + MPI) cannot be built on any box of this pool (no Fortran compiler), so this is the oracle
    port -- the restatement that tests/test_reference_pin.py pins to the reference's source text -- with
    multi-threaded pocketfft, on a bounded z-sample of the workload's plane size."""
    if rank != 0:
        return
    nx, ny, Nz = grid_of(args)
    cores = os.cpu_count() or 1
    v, sample, dt = cpu_core_step_rate(nx, ny, cores, budget_s=max(5.0, min(60.0, 4.0 * (args.steps + args.warmup))))
    line = {"impl": "reference", "metric": "grid-point updates/sec", "value": v, "unit": "Mpts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["text"] + " (CPU: bounded z-sample, core step)", "grid": [nx, ny, Nz]},
            "cpu_baseline": {"value": v, "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference cannot be built on this pool (no gfortran/FFTW3/MPI, BASELINE.md): oracle port of its "
                    "algorithm, pinned to the reference's source text by tests/test_reference_pin.py; per-step time "
                    "%.3f s on the sample" % dt}
    print(json.dumps(line), flush=True)


def parity_check(dist, rank, world, local):
    """Driver-visible multi-rank parity (outside every timed region): the `world` z-slab ranks of THIS launch advance a
    small channel two core steps through the library (halos, slab <-> pencil transposes, k = 0 chain over NCCL /
    peer memory, exactly the benchmarked code path) and rank 0 compares the gathered fields with the SINGLE-slab
    oracle.  The oracle is used here only as the checker."""
    import lesgo_b200
    from lesgo_b200 import slab
    from oracle import lesgo_oracle as O
    kw = dict(nx=64, ny=64, Nz=8 * world, lbc_mom=1, ubc_mom=1, utop=1.0, ubot=-1.0)
    pr = O.Params(nproc=world, coord=rank, **kw)
    dims = lesgo_b200.Dims(nx=pr.nx, ny=pr.ny, Nz=pr.Nz, nproc=world, coord=rank, lbc_mom=1, ubc_mom=1, sgs=False, device=local)
    core = lesgo_b200.Core(dims)
    if world > 1:
        slab.bootstrap_comm(core, dist)
    ug, vg, wg = O.synthetic_global(pr.nx, pr.ny, pr.Nz, nproc=world, seed=7, amp=0.3)
    for n, g in (("u", ug), ("v", vg), ("w", wg)):
        core.upload(n, O.scatter_slab(g, pr))
    for n in ("RHSx", "RHSy", "RHSz", "divtx", "divty", "divtz"):
        core.upload(n, np.zeros(dims.shape))
    nsteps = 2
    for it in range(nsteps):
        core.step(dt=pr.dt, tadv1=1.5, tadv2=-0.5, first_step=(it == 0), mode=0, ubot=-1.0, utop=1.0)
    names = ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz")
    mine = {n: core.download(n) for n in names}
    mine["cfl"] = core.max_cfl(pr.dt)
    alls = [mine]
    if world > 1:
        alls = [None] * world
        dist.all_gather_object(alls, mine)
    out = None
    if rank == 0:
        pg = O.Params(nproc=1, **kw)
        spg = O.Spectral(pg)
        s = O.State(pg)
        s.u, s.v, s.w = (O.scatter_slab(f, pg) for f in (ug, vg, wg))
        for it in range(nsteps):
            O.step(s, spg, O.LocalComm(), mode="core", first_step=(it == 0))
        ps = [O.Params(nproc=world, coord=r, **kw) for r in range(world)]
        worst = {}
        for n in names:
            top = n in ("w", "RHSz", "p")
            g = O.gather_slabs([alls[r][n] for r in range(world)], ps, top_extra=top)
            hi = pg.nz_tot if top else pg.nz_tot - 1
            a, b = g[1:hi + 1, :, :pg.nx], getattr(s, n)[1:hi + 1, :, :pg.nx]
            worst[n] = float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
        cfl_ref = O.get_max_cfl(s, pg, O.LocalComm())
        out = {"nranks": world, "grid": [pg.nx, pg.ny, pg.Nz], "steps": nsteps, "max_rel_l2": max(worst.values()),
               "rel_l2": worst, "max_cfl_rel_err": max(abs(alls[r]["cfl"] - cfl_ref) / cfl_ref for r in range(world)),
               "gate": 1e-12, "pass": bool(max(worst.values()) <= 1e-12),
               "checker": "single-slab oracle (oracle/lesgo_oracle.py, pinned to the reference sources by "
                          "tests/test_reference_pin.py); transposes: " +
                          ("n/a" if world == 1 else ("NVLink peer memory" if getattr(core, "p2p_enabled", False) else "NCCL"))}
    del core
    return out


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
    wl = WORKLOADS[args.workload]
    kind = wl["kind"] if not args.grid else "core"
    nx, ny, Nz = grid_of(args)

    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(dist, rank, world, local)
        except Exception as e:  # noqa
            parity = {"error": str(e)}
        if rank == 0 and parity and parity.get("pass") is False:
            raise SystemExit(f"bench.py: {world}-rank parity check failed: {parity}")

    wall = dict(lbc_mom=1, ubc_mom=1) if kind == "core" else dict(lbc_mom=2, ubc_mom=0)
    dims = lesgo_b200.Dims(nx=nx, ny=ny, Nz=Nz, nproc=world, coord=rank, sgs=True, device=local, **wall)
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
    farm = None
    if kind == "lasd":
        for n in ("F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2"):
            core.upload(n, zero)
        step_kw = dict(step_kw, mode=1, sgs_model=5, nu=1e-4, lagran_dt=5 * dt)
    elif kind == "adm":
        farm = synthetic_farm(dims)
        core.turbines_init(farm)
        step_kw = dict(step_kw, mode=1, sgs_model=1, nu=1e-4, turbines=True, turbines_eps=0.1)
    del zero
    jt = [0]

    def one_step():
        """One timestep of the workload; the LASD workload runs lagrange_Sdep on every fifth step (cs_count = 5,
        sgs_stag_util.f90:192), as the shipped LES_channel_Re1000 does."""
        jt[0] += 1
        if kind == "lasd":
            if jt[0] == 1:
                core.step(first_step=True, lasd_cs_init=True, **step_kw)
            elif jt[0] == 2:
                core.step(lasd_update=True, lasd_init_F=True, **step_kw)
            else:
                core.step(lasd_update=(jt[0] % 5 == 0), **step_kw)
        else:
            core.step(first_step=(jt[0] == 1), **step_kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nwarm = max(args.warmup, 3)
    if kind == "lasd":
        nwarm = max(nwarm, 5)
        while (nwarm % 5) != 0:
            nwarm += 1                       # the timed region starts at a multiple of cs_count
    for _ in range(nwarm):
        one_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = core.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    for i in range(args.steps):
        ev[i].record(stream)
        one_step()
    ev[args.steps].record(stream)
    barrier()
    ms = ev[0].elapsed_time(ev[args.steps])
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    launches = core.launch_count - l0
    clocks = sampler.stop() if sampler else None
    ms = allmax(ms)
    cfl = core.max_cfl(dt)
    if not math.isfinite(cfl) or cfl <= 0.0 or cfl > 10.0:
        raise SystemExit(f"bench.py: simulation state is not sane after the timed steps (CFL = {cfl})")
    ms_step = ms / args.steps
    med = allmax(per_step[len(per_step) // 2])
    points = nx * ny * (dims.nz_tot - 1)
    value = points / (ms_step * 1e-3) / 1e6

    # per-pass breakdown: one instrumented step outside the timed region
    core.profile(True)
    one_step()
    kern = core.profile(False, report=True)
    peak, peak_src = hbm_peak()
    abytes = A_CORE_BYTES if kind == "core" else A_STEP_BYTES
    achieved = abytes * points / (ms_step * 1e-3) / 1e9 / world
    traffic, traffic_src = None, None
    try:
        for tname in ("r5_traffic.json", "r4_traffic.json", "r3_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                if tj.get("grid") == [nx, ny, Nz] and world == 1 and kind == "core":
                    traffic = tj["dram_bytes_per_step"]
                    traffic_src = "FROM FILE profiles/%s, not measured in this run: %s" % (tname, tj["source"])
                break
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": abytes * points,
                "peak_source": peak_src,
                "definition": ("296 B/point/step (SURVEY 8d A_core)" if kind == "core" else
                               "976 B/point/step (SURVEY 8d A_step, the whole device-resident step in the reference's dataflow)")
                + " * points / step time / n_gpus; whole hot path, all kernels of the step"}
    kernels = {k: {"launches": n, "ms": round(t, 4)} for k, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1])}

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
        return allmax(a.elapsed_time(b) / n)

    # the complete step and rows (f)-2 ... (f)-4 of SURVEY section 8 on the headline grid, timed separately
    full = lasd = turb = tavg = None
    if kind == "core" and not args.no_extras:
        try:
            fms = timed_steps(dict(step_kw, mode=1, sgs_model=1, nu=1e-4), max(2, args.steps // 4))
            full = {"ms_per_step": fms, "value": points / (fms * 1e-3) / 1e6, "unit": "Mpts/s",
                    "what": "core + DNS-wall wallstress + calc_Sij + Smagorinsky sgs_stag + divstress_uv/w"}
        except Exception as e:  # noqa
            full = {"error": str(e)}
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
        try:
            farm = synthetic_farm(dims)
            core.turbines_init(farm)
            t_ms = timed_steps(dict(step_kw, mode=1, sgs_model=1, nu=1e-4, turbines=True, turbines_eps=0.1), 3)
            turb = {"ms_per_step": t_ms, "value": points / (t_ms * 1e-3) / 1e6, "unit": "Mpts/s", "disks": len(farm),
                    "nodes_this_rank": int(sum(len(t.ind) for t in farm)),
                    "what": "full step (Smagorinsky) + turbines_forcing: gather, all-reduce, scatter, RHS += f"}
        except Exception as e:  # noqa
            turb = {"error": str(e)}
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
    if not args.no_e2e:
        e2e = run_e2e(core, dims, u, v, w, dt, tadv1, args.e2e_steps, points, kind, step_kw, barrier, allmax, world)
    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        cv, sample, _ = cpu_core_step_rate(nx, ny, cores, budget_s=15.0)
        cpu = {"value": cv, "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        line = {"metric": "grid-point updates/sec", "value": value, "unit": "Mpts/s", "n_gpus": world,
                "steps": args.steps, "warmup": nwarm, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["text"] if not args.grid else f"LES channel core timestep {nx}x{ny}x{Nz} FP64 (device-resident)",
                           "name": args.workload if not args.grid else "custom-grid core step",
                           "grid": [nx, ny, Nz], "decomposition": f"z-slabs x{world}",
                           "pressure_transposes": ("n/a" if world == 1 else
                                                   ("NVLink peer-memory stores" if getattr(core, "p2p_enabled", False)
                                                    else "NCCL all-to-all")),
                           "l2": "inputs larger than L2 (%.0f MB per field)" % (np.prod(dims.shape) * 8 / 1e6)},
                "per_step_ms": {"median": med, "min": per_step[0], "max": per_step[-1],
                                "how": "one CUDA event pair per step inside the same timed region (rank 0; median = max over ranks)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "parity": parity,
                "clocks": clocks, "kernels": kernels, "max_cfl": cfl, "full_step": full, "lasd_step": lasd, "turbines_step": turb, "tavg": tavg}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(core, dims, u, v, w, dt, tadv1, nsteps, points, kind, step_kw, barrier, allmax, world):
    """The same metric end to end through the reference-facing C ABI with HOST buffers, every rank on its own slab,
    host<->device copies inside the timed region (wall clock around synchronised calls, max over ranks).

    core workload -- the per-routine entry points the Fortran shims bind (filt_da x3, ddz_uv x2, ddz_w, convec,
    press_stag_array), each staging its inputs H2D and its outputs D2H, measured three ways: the arrays page-locked
    with lesgo_gpu_host_register as fortran/lesgo_gpu_mod.f90 does for the module arrays (the headline `value`),
    plain pageable arrays (what an unmodified host would pass), and for reference torch-pinned arrays.
    Every workload -- `whole_step_api`: the resident-state entry (upload u, v, w, RHS*; lesgo_gpu_step; download the
    seven result fields), which is what the shims' resident module uses and the headline for the full-step workloads.
    """
    import torch

    def host_arrays(mode, names, init):
        out = {}
        for n in names:
            if mode == "pinned":
                t = torch.empty(dims.shape, dtype=torch.float64, pin_memory=True)
                a = t.numpy()
                out["_t_" + n] = t                   # keeps the pinned allocation alive
            else:
                a = np.empty(dims.shape)
            a[...] = init.get(n, 0.0)
            if mode == "registered":
                core.host_register(a)
            out[n] = a
        return out

    nb = float(np.prod(dims.shape) * 8)
    res = {}
    if kind == "core":
        names = ("dudx", "dudy", "dudz", "dvdx", "dvdy", "dvdz", "dwdx", "dwdy", "dwdz", "RHSx", "RHSy", "RHSz", "divtz",
                 "p", "dpdx", "dpdy", "dpdz", "u", "v", "w")
        h2d = nb * (3 * 1 + 3 * 1 + 9 + 4)          # filt_da in; ddz in; convec in; press in (outputs are not uploaded)
        d2h = nb * (3 * 3 + 3 + 3 + 4)
        variants = {}
        for mode in ("registered", "pageable", "pinned"):
            F = host_arrays(mode, names, {"u": u, "v": v, "w": w})

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
            barrier()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                one()
            barrier()
            t = allmax((time.perf_counter() - t0) / nsteps)
            variants[mode] = {"value": points / t / 1e6, "ms_per_step": t * 1e3}
            if mode == "registered":
                for n in names:
                    core.host_unregister(F[n])
            del F
        res = {"value": variants["registered"]["value"], "unit": "Mpts/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": variants["registered"]["ms_per_step"],
               "api": "per-routine C ABI (filt_da x3, ddz_uv x2, ddz_w, convec, press_stag_array) on host arrays page-locked "
                      "with lesgo_gpu_host_register, as fortran/lesgo_gpu_mod.f90 registers the sim_param arrays",
               "host_memory": variants, "n_gpus": world}
    # the whole-step entry with the state held by the host: upload u, v, w, RHSx, RHSy, RHSz, one lesgo_gpu_step,
    # download u, v, w, p, RHSx, RHSy, RHSz, every step
    try:
        S = host_arrays("registered", ("RHSx", "RHSy", "RHSz", "p", "u", "v", "w"), {"u": u, "v": v, "w": w})
        kw = dict(step_kw)
        if kind == "lasd":
            kw["lasd_update"] = False

        def one_step(first=False):
            for n in ("u", "v", "w", "RHSx", "RHSy", "RHSz"):
                core.upload(n, S[n])
            core.step(first_step=first, **kw)
            for n in ("u", "v", "w", "p", "RHSx", "RHSy", "RHSz"):
                core.download(n, S[n])

        one_step(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            one_step()
        barrier()
        ts = allmax((time.perf_counter() - t0) / nsteps)
        ws = {"value": points / ts / 1e6, "unit": "Mpts/s", "ms_per_step": ts * 1e3,
              "h2d_bytes_per_step": 6 * nb * world, "d2h_bytes_per_step": 7 * nb * world,
              "api": "lesgo_gpu_upload x6 + lesgo_gpu_step + lesgo_gpu_download x7, host arrays page-locked with "
                     "lesgo_gpu_host_register (fortran/lesgo_gpu_resident_mod.f90)"}
        for n in ("RHSx", "RHSy", "RHSz", "p", "u", "v", "w"):
            core.host_unregister(S[n])
    except Exception as e:  # noqa
        ws = {"error": str(e)}
    if kind == "core":
        res["whole_step_api"] = ws
    else:
        res = dict(ws)
        res["n_gpus"] = world
    return res


if __name__ == "__main__":
    main()
