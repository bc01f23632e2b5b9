"""refrun -- drive the reference's OWN Fortran sources (under /root/reference) through oracle/f90exec.py.

TEST INFRASTRUCTURE ONLY.  This is the closest thing to "the reference itself run here" that a container
without any Fortran compiler allows (probes: BASELINE.md): the statements of fft.f90, emul_complex.f90,
derivatives.f90, convec.f90, press_stag_array.f90, tridag_array.f90, forcing.f90 (project), cfl_util.f90,
mpi_defs.f90, wallstress.f90, sgs_stag_util.f90, divstress_uv/w.f90, sim_param.f90 and the time-loop body of
main.f90 are interpreted as they stand, with the reference's default build flags (USE_MPI, USE_SAFETYMODE;
CMakeLists.txt:22-31), one rank (an MPI build run with -np 1: lbz = 0, neighbours MPI_PROC_NULL).  Bound by this
file, because they live outside the reference's sources:
  * FFTW3 (dfftw_plan_dft_*_2d, dfftw_execute_dft_r2c / c2r): pocketfft through the oracle's r2c / c2r, so that
    a difference between this run and the oracle isolates the restated LOGIC, not FFT rounding;
  * MPI (mpi_sendrecv / send / recv to MPI_PROC_NULL are no-ops, mpi_allreduce over one rank is a copy);
  * lesgo.conf parsing (input_util.f90): the derived sizes of input_util.f90:197-235 are set from Params.
Nothing is copied: the reference text is read at run time.  /root/reference does not exist on the GPU boxes, so
only oracle/make_reference_fixtures.py (run here) imports this; tests use the frozen tests/golden/ref_*.npz.
"""
from __future__ import annotations

import os

import numpy as np

from . import f90exec as F
from . import lesgo_oracle as O

REF = os.environ.get("LESGO_REFERENCE_DIR", "/root/reference")
FILES = ["types.f90", "param.f90", "sim_param.f90", "messages.f90", "emul_complex.f90", "fft.f90", "derivatives.f90",
         "convec.f90", "tridag_array.f90", "press_stag_array.f90", "cfl_util.f90", "forcing.f90", "mpi_defs.f90",
         "sgs_param.f90", "test_filtermodule.f90", "wallstress.f90", "sgs_stag_util.f90", "divstress_uv.f90",
         "divstress_w.f90"]
# + the Lagrangian scale-dependent dynamic model (rows (f)-2): grid_m (derived type with pointer components),
# trilinear_interp_w / cell_indx in functions.f90, lagrange_Sdep.f90, interpolag_Sdep.f90
LASD_FILES = FILES + ["grid.f90", "functions.f90", "lagrange_Sdep.f90", "interpolag_Sdep.f90"]
# + the running time averages (rows (f)-4): tavg%compute of time_average.f90 (a derived type with 26 allocatable components)
TAVG_FILES = LASD_FILES + ["stat_defs.f90", "time_average.f90"]
# + actuator disks (rows (f)-3): turbines_forcing of turbines.f90 (arrays of derived types, scalar pointers), needs the
# PPTURBINES build flag (USE_TURBINES, CMakeLists.txt:27) for sim_param's fxa, fya, fza and forcing_applied
TURBINE_FILES = ["types.f90", "param.f90", "sim_param.f90", "messages.f90", "emul_complex.f90", "fft.f90", "derivatives.f90",
                 "convec.f90", "tridag_array.f90", "press_stag_array.f90", "cfl_util.f90", "mpi_defs.f90", "sgs_param.f90",
                 "test_filtermodule.f90", "wallstress.f90", "sgs_stag_util.f90", "divstress_uv.f90", "divstress_w.f90",
                 "grid.f90", "functions.f90", "turbine_indicator.f90", "stat_defs.f90", "turbines.f90", "forcing.f90"]
MPI_PROC_NULL = -2


def available():
    return os.path.isdir(REF) and os.path.exists(os.path.join(REF, "convec.f90"))


class Reference:
    """One rank of the reference, interpreted.  Fields are the module arrays of sim_param (Fortran bounds kept)."""

    def __init__(self, p: O.Params, files=FILES, alloc_fill=0.0, dyn_init=100, cs_count=5, turbines=False, boxes=None,
                 overrides=None, post_load=None):
        """boxes: shared mailbox dict of a multi-rank run (run_ranks below); None = one rank.
        overrides: {file name: path} -- sources taken from elsewhere than the reference tree (an absolute path in `files`
        is loaded as it is); post_load(self): called after the sources are loaded and the externals installed, before
        any reference code runs.  Both exist for tests/test_shim_dropin.py, which swaps the five replaced sources for
        the ISO_C_BINDING shims of fortran/."""
        assert p.nproc == 1 or boxes is not None, "several ranks need the shared mailboxes of run_ranks()"
        self.p = p
        self.boxes = boxes
        self.turbines = turbines
        I = self.I = F.Interpreter(defines=("PPMPI", "PPSAFETYMODE") + (("PPTURBINES",) if turbines else ()), alloc_fill=alloc_fill)
        for f in files:
            I.load((overrides or {}).get(f) or (f if os.path.isabs(f) else os.path.join(REF, f)))
        self.plans = {}
        self._externals()
        if post_load is not None:
            post_load(self)
        S = lambda n, v: I.set("param", n, v)
        # input_util.f90:197-235 (derived sizes) and the lesgo.conf blocks this path reads
        S("nproc", p.nproc); S("coord", p.coord); S("rank", p.coord)
        S("nx", p.nx); S("ny", p.ny); S("nz", p.nz); S("nz_tot", p.nz_tot)
        S("nx2", p.nx2); S("ny2", p.ny2); S("lh", p.lh); S("ld", p.ld); S("lh_big", p.lh_big); S("ld_big", p.ld_big)
        S("l_x", p.L_x); S("l_y", p.L_y); S("l_z", p.L_z); S("z_i", p.z_i)
        S("dx", p.dx); S("dy", p.dy); S("dz", p.dz)
        S("dt", p.dt); S("tadv1", p.tadv1); S("tadv2", p.tadv2); S("dt_f", p.dt)
        S("lbc_mom", p.lbc_mom); S("ubc_mom", p.ubc_mom); S("sgs", bool(p.sgs)); S("molec", bool(p.molec))
        S("sgs_model", p.sgs_model); S("nu_molec", p.nu_molec); S("u_star", p.u_star); S("co", p.Co)
        S("wall_damp_exp", p.wall_damp_exp); S("vonk", p.vonk); S("zo", p.zo); S("ifilter", p.ifilter)
        S("ubot", p.ubot); S("utop", p.utop)
        S("use_mean_p_force", bool(p.use_mean_p_force)); S("mean_p_force_x", p.mean_p_force_x); S("mean_p_force_y", p.mean_p_force_y)
        # mpi_defs.f90:77-87: 1-D chain, MPI_PROC_NULL beyond the ends
        S("up", p.coord + 1 if p.coord + 1 < p.nproc else MPI_PROC_NULL)
        S("down", p.coord - 1 if p.coord > 0 else MPI_PROC_NULL); S("comm", 0); S("ierr", 0); S("mpi_rprec", 0)
        S("status", F.FArray.alloc((8,), (1,), "integer"))
        S("initu", False); S("jt_total", 0); S("jt", 0); S("use_cfl_dt", False); S("cfl", 0.0625)
        S("inilag", True); S("dyn_init", int(dyn_init)); S("cs_count", int(cs_count))
        I.call("sim_param_init", module="sim_param")              # sim_param.f90: allocates the 33 module arrays
        if "sgs_param" in I.modules and "sgs_param_init" in I.modules["sgs_param"].procs:
            I.call("sgs_param_init", module="sgs_param")          # initialize.f90:129
        I.call("init_fft", module="fft")                          # initialize.f90:172; fft.f90:102-160: plans + wavenumbers
        if "grid_m" in I.modules and "build" in I.modules["grid_m"].procs:
            I.call("build", I.get("grid_m", "grid"), module="grid_m")     # initialize.f90:132 call grid%build()
        if "test_filtermodule" in I.modules and "test_filter_init" in I.modules["test_filtermodule"].procs:
            I.call("test_filter_init", module="test_filtermodule")   # initialize.f90:176

    # ---- what lives outside the reference's own sources ---------------------------------------------
    def _externals(self):
        I = self.I
        plans = self.plans

        def plan(kind):
            def f(fr, a):
                h = len(plans) + 1
                plans[h] = (kind, int(a[1][0]), int(a[2][0]))
                a[0][1](h)                                           # integer*8 plan handle, by reference
            return f

        def execute(kind):
            def f(fr, a):
                k, n0, n1 = plans[int(a[0][0])]
                assert k == kind, "plan direction"
                src, dst = a[1][0], a[2][0]
                x = src.a.T                                          # (n_slow, 2*(n_fast/2+1)), C order view
                assert x.shape == (n1, 2 * (n0 // 2 + 1)), (x.shape, n0, n1)
                y = O.r2c(np.ascontiguousarray(x)[None], n0)[0] if kind == "r2c" else O.c2r(np.ascontiguousarray(x)[None], n0)[0]
                dst.a.T[...] = y
            return f

        me, nproc, boxes = self.p.coord, self.p.nproc, self.boxes

        def flat(x, count):
            """the `count` storage units an MPI buffer argument designates: a section (view) or the sequence that starts
            at an array element (sequence association, e.g. rH_x(1, 1, nz-1))"""
            if isinstance(x, F.ElemRef):
                f_ = x.base.a.reshape(-1, order="F")
                assert np.shares_memory(f_, x.base.a)
                start = int(np.ravel_multi_index(x.idx, x.base.a.shape, order="F"))
                return f_[start:start + count]
            f_ = x.a.reshape(-1, order="F")
            if not np.shares_memory(f_, x.a):
                # a non-contiguous section, e.g. var(:, :, 1) of fxa(1:nx, 1:ny, lbz:nz): the compiler passes a
                # contiguous temporary (copy-in / copy-out); the whole section is the message
                assert count == x.a.size, "a non-contiguous MPI buffer must be sent or received whole"
                return None
            return f_[:count]

        def box(src, dst, tag):
            import queue
            with boxes["lock"]:
                return boxes.setdefault((src, dst, tag), queue.Queue())

        def valid(r):
            return r != MPI_PROC_NULL and 0 <= r < nproc

        def send_to(buf, count, dest, tag):
            if valid(dest):
                v = flat(buf, count)
                box(me, dest, tag).put(np.array(v, copy=True) if v is not None else buf.a.flatten(order="F"))

        def recv_from(buf, count, src, tag):
            if valid(src):
                data = box(src, me, tag).get(timeout=MPI_TIMEOUT)
                v = flat(buf, count)
                if v is not None:
                    v[...] = data
                else:
                    buf.a[...] = np.asarray(data).reshape(buf.a.shape, order="F")

        def sendrecv(fr, a):
            send_to(a[0][0], int(a[1][0]), int(a[3][0]), int(a[4][0]))
            recv_from(a[5][0], int(a[6][0]), int(a[8][0]), int(a[9][0]))

        def send(fr, a):
            send_to(a[0][0], int(a[1][0]), int(a[3][0]), int(a[4][0]))

        def recv(fr, a):
            recv_from(a[0][0], int(a[1][0]), int(a[3][0]), int(a[4][0]))

        def allreduce(fr, a):
            op = a[4][0][1] if isinstance(a[4][0], tuple) else "mpi_sum"
            mine = np.array(a[0][0].a, copy=True) if isinstance(a[0][0], F.FArray) else np.array(a[0][0])
            acc = None
            for r in range(nproc):
                if r != me:
                    box(me, r, ("ar", op)).put(mine)
            for r in range(nproc):                                # rank order: every rank forms the same result
                x = mine if r == me else box(r, me, ("ar", op)).get(timeout=MPI_TIMEOUT)
                acc = x if acc is None else (np.maximum(acc, x) if op == "mpi_max" else
                                             (np.minimum(acc, x) if op == "mpi_min" else acc + x))
            if isinstance(a[1][0], F.FArray):
                a[1][0].a[...] = acc
            else:
                a[1][1](float(acc))

        def reduce_(fr, a):
            # mpi_reduce(send, recv, count, type, op, root, comm, ierr): sum in rank order on the root only
            root = int(a[5][0])
            mine = np.array(a[0][0].a, copy=True) if isinstance(a[0][0], F.FArray) else np.array(a[0][0])
            if me != root:
                box(me, root, ("red",)).put(mine)
                return
            acc = None
            for r in range(nproc):
                x = mine if r == me else box(r, me, ("red",)).get(timeout=MPI_TIMEOUT)
                acc = x if acc is None else acc + x
            if isinstance(a[1][0], F.FArray):
                a[1][0].a[...] = acc
            else:
                a[1][1](float(acc))

        def error(fr, a):
            raise F.FortranError("reference called error(): " + " ".join(str(x[0]) for x in a))

        I.externals.update({
            "dfftw_plan_dft_r2c_2d": plan("r2c"), "dfftw_plan_dft_c2r_2d": plan("c2r"),
            "dfftw_execute_dft_r2c": execute("r2c"), "dfftw_execute_dft_c2r": execute("c2r"),
            "mpi_sendrecv": sendrecv, "mpi_send": send, "mpi_recv": recv, "mpi_allreduce": allreduce, "mpi_reduce": reduce_,
            "error": error, "apply_inflow": lambda fr, a: None, "mpi_barrier": lambda fr, a: None,
        })

    # ---- fields ---------------------------------------------------------------------------------------
    def farray(self, name):
        return self.I.get("sim_param", name)

    def get(self, name, module="sim_param"):
        """module array `name` as the oracle lays fields out: (0:nz, ny, ld); arrays declared 1:nz get a zero plane 0."""
        fa = self.I.get(module, name)
        out = np.zeros((self.p.nz + 1, self.p.ny, self.p.ld))
        k0 = fa.lb[2]
        out[k0:k0 + fa.a.shape[2]] = fa.a.transpose(2, 1, 0)
        return out

    def put(self, name, arr):
        fa = self.farray(name)
        k0 = fa.lb[2]
        fa.a[...] = np.asarray(arr)[k0:k0 + fa.a.shape[2]].transpose(2, 1, 0)

    def farm_set(self, farm, eps, adm_correction=False, use_rotation=False, tip_speed_ratio=7.0):
        """wind_farm as turbines_init / turbines_nodes leave it (turbines.f90:129-462 are host start-up work fed from input
        files; both sides are handed the same node lists): nloc disks with %nodes, %ind, %nhat, %Ct_prime, %dia, %u_d_T,
        %turb_ind_func%M.  eps enters through T_avg_dim and dt_dim exactly as turbines.f90:563-567 forms it."""
        I, p = self.I, self.p
        sd = I.modules["stat_defs"].types
        wf = I.get("stat_defs", "wind_farm")
        arr = F.FArray.alloc((len(farm),), (1,), "object")
        for s_, t in enumerate(farm):
            o = F.FStruct(sd["turbine_t"])
            n = len(t.ind)
            o.num_nodes = n
            o.nodes = F.FArray(np.asfortranarray(np.asarray(t.nodes, dtype=np.int64).reshape(n, 3)), (1, 1), "integer")
            o.ind = F.FArray(np.array(t.ind, dtype=np.float64), (1,))
            o.nhat = F.FArray(np.array(t.nhat, dtype=np.float64), (1,))
            if use_rotation:                                          # %ind_t, %e_theta: turbines.f90:419-429, :456
                o.ind_t = F.FArray(np.array(t.ind_t, dtype=np.float64), (1,))
                o.e_theta = F.FArray(np.asfortranarray(np.asarray(t.e_theta, dtype=np.float64).reshape(n, 3)), (1, 1))
            o.ct_prime, o.dia, o.u_d_t, o.u_d, o.f_n = float(t.Ct_prime), float(t.dia), float(t.u_d_T), 0.0, 0.0
            o.theta1, o.theta2 = float(getattr(t, "theta1", 0.0)), float(getattr(t, "theta2", 0.0))
            o.icp = o.jcp = o.kcp = 1
            o.center_in_proc = False
            tif = F.FStruct(I.modules["turbine_indicator"].types["turb_ind_func_t"])
            tif.m = float(getattr(t, "M", 1.0))
            o.turb_ind_func = tif
            arr.a[s_] = o
        wf.turbine = arr
        T = lambda n, v: I.set("turbines", n, v)
        T("nloc", len(farm)); T("dyn_theta1", False); T("dyn_theta2", False); T("dyn_ct_prime", False)
        T("adm_correction", bool(adm_correction)); T("use_rotation", bool(use_rotation)); T("tbase", 10 ** 9)
        T("tip_speed_ratio", float(tip_speed_ratio))
        # eps = (dt_dim / T_avg_dim) / (1 + dt_dim / T_avg_dim)  ->  dt_dim / T_avg_dim = eps / (1 - eps)
        I.set("param", "dt_dim", float(eps / (1.0 - eps))); T("t_avg_dim", 1.0)
        I.set("param", "total_time_dim", 0.0); I.set("param", "total_time", 0.0)
        T("vel_top_dat", "vel_top.dat"); T("forcing_fid", F.FArray.alloc((len(farm),), (1,), "integer"))
        self.farm_objs = arr

    def farm_get(self, name):
        return [getattr(o, name) for o in self.farm_objs.a]

    def tavg_new(self):
        """A tavg_t as tavg%init leaves it (time_average.f90:77-170 minus the file I/O): zeroed accumulators
        (nx, ny, lbz:nz), total_time = 0, plus the module's work arrays w_uv ... vortz."""
        p, I = self.p, self.I
        t = F.FStruct(I.modules["time_average"].types["tavg_t"])
        mk = lambda: F.FArray.alloc((p.nx, p.ny, p.nz + 1), (1, 1, 0))
        for n, d in t._type.members.items():
            if d.dims:
                setattr(t, n, mk())
        t.total_time, t.dt, t.initialized = 0.0, 0.0, True
        for n in ("w_uv", "u_w", "v_w", "vortx", "vorty", "vortz", "pres_real"):
            I.set("time_average", n, mk())
        return t

    def tavg_compute(self, t, dt):
        t.dt = float(dt)                                   # io.f90 output_loop sets tavg%dt before tavg%compute
        self.I.call("compute", t, module="time_average")

    def set_dt(self, dt, tadv1, tadv2):
        self.I.set("param", "dt", float(dt)); self.I.set("param", "tadv1", float(tadv1)); self.I.set("param", "tadv2", float(tadv2))

    # ---- routines -------------------------------------------------------------------------------------
    def call(self, name, *args, module=None):
        return self.I.call(name, *args, module=module)

    def step(self, jt_total, mode="core"):
        """One pass through the time-loop body of main.f90 (lines cited are the reference's): RHS_f = RHS, filt_da,
        ddz, wallstress, [sgs_stag, tzz halo, divstress], convec, RHS assembly, mean pressure forcing, Euler start,
        AB2, BOGUS, press_stag_array, RHS -= grad p, project.  mode "core" skips :189-203 (divt* stay as they are),
        like the oracle's core mode."""
        I = self.I
        main = os.path.join(REF, "main.f90")
        uses = ["types", "param", "sim_param", "derivatives", "forcing", "fft", "sgs_stag_util", "cfl_util"]
        I.set("param", "jt_total", int(jt_total)); I.set("param", "jt", int(jt_total))     # main.f90:147-148, fresh run
        I.exec_lines(main, 155, 184, uses)
        if mode == "full":
            I.exec_lines(main, 189, 203, uses)
        I.exec_lines(main, 207, 214, uses)
        I.exec_lines(main, 229, 232, uses)
        if self.turbines:
            I.exec_lines(main, 254, 254, uses)                      # call forcing_applied() -> turbines_forcing
            I.exec_lines(main, 263, 267, uses)                      # RHS += fxa, fya, fza
        I.exec_lines(main, 273, 326, uses)
        I.exec_lines(main, 344, 344, uses)

    def cfl_dt_step_setup(self):
        """main.f90:135-144 (use_cfl_dt): executed from the reference text."""
        main = os.path.join(REF, "main.f90")
        self.I.set("param", "use_cfl_dt", True)
        self.I.exec_lines(main, 135, 144, ["types", "param", "sim_param", "cfl_util"], local={"dt_dim": 0.0})


MPI_TIMEOUT = float(os.environ.get("REFRUN_MPI_TIMEOUT", "600"))   # seconds a blocking receive waits for its message


def run_ranks(kw, nproc, fn, **ref_kw):
    """The reference on `nproc` ranks: one interpreter per rank, each in its own thread, MPI calls carried by in-process
    mailboxes (blocking, tag-matched, like the MPI the reference uses).  fn(ref, coord) runs on every rank; returns the
    list of its results in rank order."""
    import threading
    boxes = {"lock": threading.Lock()}
    res, err = [None] * nproc, [None] * nproc

    def work(r):
        try:
            ref = Reference(O.Params(nproc=nproc, coord=r, **kw), boxes=boxes, **ref_kw)
            res[r] = fn(ref, r)
        except BaseException as e:  # noqa
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nproc)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    import queue
    real = [e for e in err if e is not None and not isinstance(e, queue.Empty)]
    if real:
        raise real[0]              # a rank that failed, not the peers that then waited for it in vain
    for e in err:
        if e is not None:
            raise e
    return res
