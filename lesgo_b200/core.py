"""Host-side mirror of the reference interface for the hot path.

`Core` owns one library context (= one MPI rank / one GPU of the reference's z-slab
decomposition, mpi_defs.f90:77-87) and exposes the reference's subroutine names with
the same argument meaning:

    derivatives.f90 : ddx, ddy, ddxy, filt_da, ddz_uv, ddz_w
    convec.f90      : convec
    press_stag_array.f90 : press_stag_array          (tridag_array inside)
    fft.f90         : padd, unpadd, wavenumbers, fft_r2c / fft_c2r (the FFTW plans)
    main.f90 loop   : step  (device-resident fields), max_cfl, rmsdiv

Arrays are the reference's `(ld, ny, 0:nz)` Fortran arrays held as C-ordered
`[k, j, i]` arrays of shape `(nz+1, ny, ld)`: numpy arrays (host; staged through the
device inside each call, like the Fortran shim does) or CUDA `torch.float64` tensors
(used in place).  Errors raise `LibraryError` with the library's message, the analogue
of the reference's `call error(...)` (messages.f90:228-240).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from .lib import DimsStruct, Library, LibraryError, StepParams, TurbineStruct, load_library

FIELD_IDS = {n: i for i, n in enumerate(
    ["u", "v", "w", "dudx", "dudy", "dudz", "dvdx", "dvdy", "dvdz", "dwdx", "dwdy", "dwdz",
     "RHSx", "RHSy", "RHSz", "RHSx_f", "RHSy_f", "RHSz_f", "p", "dpdx", "dpdy", "dpdz",
     "divtx", "divty", "divtz", "txx", "txy", "txz", "tyy", "tyz", "tzz",
     "F_LM", "F_MM", "F_QN", "F_NN", "Cs_opt2", "fxa", "fya", "fza"])}


TAVG_IDS = {n: i for i, n in enumerate(
    ["u", "v", "w", "w_uv", "u_w", "v_w", "u2", "v2", "w2", "uv", "uw", "vw", "txx", "tyy", "tzz", "txy", "txz", "tyz",
     "p", "fx", "fy", "fz", "cs_opt2", "vortx", "vorty", "vortz"])}


@dataclass
class Dims:
    """Grid / decomposition parameters, named as in param.f90 / lesgo.conf."""
    nx: int
    ny: int
    Nz: int                      # lesgo.conf "Nz"; per-rank nz = Nz//nproc + 1 (input_util.f90:197)
    nproc: int = 1
    coord: int = 0
    L_x: float = 2.0 * math.pi
    L_y: float = 2.0 * math.pi
    L_z: float = 2.0
    lbc_mom: int = 1
    ubc_mom: int = 1
    sgs: bool = False
    device: int = -1

    @property
    def nz(self):
        return self.Nz // self.nproc + 1

    @property
    def nz_tot(self):
        return (self.nz - 1) * self.nproc + 1

    @property
    def dz(self):
        return self.L_z / (self.nz_tot - 1)

    @property
    def ld(self):
        return 2 * (self.nx // 2 + 1)

    @property
    def lh(self):
        return self.nx // 2 + 1

    @property
    def shape(self):
        return (self.nz + 1, self.ny, self.ld)

    @property
    def shape_big(self):
        return (self.nz + 1, 3 * self.ny // 2, 2 * (3 * self.nx // 4 + 1))


def _addr(a, shape=None, writable=False):
    """Address of a numpy array or torch tensor holding C-contiguous float64 data."""
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 or not a.flags.c_contiguous:
            raise LibraryError("arrays must be C-contiguous float64")
        if writable and not a.flags.writeable:
            raise LibraryError("output array is read-only")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise LibraryError(f"array shape {a.shape} != expected {tuple(shape)}")
        return a.ctypes.data
    # torch tensor (duck-typed so numpy-only callers never import torch)
    if hasattr(a, "data_ptr"):
        import torch
        if a.dtype != torch.float64 or not a.is_contiguous():
            raise LibraryError("tensors must be contiguous float64")
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise LibraryError(f"tensor shape {tuple(a.shape)} != expected {tuple(shape)}")
        return a.data_ptr()
    raise LibraryError(f"unsupported array type {type(a)}")


class Core:
    def __init__(self, dims: Dims, lib: Library | None = None):
        self.lib = lib or load_library()
        self.dims = dims
        d = DimsStruct(dims.nx, dims.ny, dims.nz, dims.nz_tot, dims.nproc, dims.coord,
                       dims.L_x, dims.L_y, dims.dz, dims.lbc_mom, dims.ubc_mom, int(dims.sgs), dims.device)
        self._ctx = C.c_void_p()
        if self.lib.create(C.byref(d), C.byref(self._ctx)):
            raise LibraryError("lesgo_gpu_create: " + self.lib.error(None))

    # -- plumbing ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise LibraryError(f"{what}: {self.lib.error(self._ctx)}")

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.set_stream(self._ctx, C.c_void_p(cuda_stream)), "set_stream")

    def synchronize(self):
        self._ck(self.lib.synchronize(self._ctx), "synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.launch_count(self._ctx))

    def profile(self, enable=True, report=False):
        """Per-launch CUDA-event timing; returns {label: (count, total_ms)} when report=True."""
        buf = C.create_string_buffer(8192) if report else None
        self._ck(self.lib.profile(self._ctx, int(enable), buf, 8192 if report else 0), "profile")
        out = {}
        if report:
            for line in buf.value.decode().splitlines():
                lab, n, ms = line.split()
                out[lab] = (int(n), float(ms))
        return out

    def empty(self, big=False):
        return np.zeros(self.dims.shape_big if big else self.dims.shape)

    # -- module fft ---------------------------------------------------------------------
    def wavenumbers(self):
        d = self.dims
        kx, ky, k2 = (np.zeros((d.ny, d.lh)) for _ in range(3))
        self._ck(self.lib.wavenumbers(self._ctx, kx.ctypes.data, ky.ctypes.data, k2.ctypes.data), "wavenumbers")
        return kx, ky, k2

    def fft_r2c(self, a, out=None, big=False):
        out = a if out is None else out
        self._ck(self.lib.fft_r2c(self._ctx, _addr(a), _addr(out, writable=True), a.shape[0], int(big)), "fft_r2c")
        return out

    def fft_c2r(self, a, out=None, big=False):
        out = a if out is None else out
        self._ck(self.lib.fft_c2r(self._ctx, _addr(a), _addr(out, writable=True), a.shape[0], int(big)), "fft_c2r")
        return out

    def padd(self, u_big, u):
        self._ck(self.lib.padd(self._ctx, _addr(u_big, writable=True), _addr(u), u.shape[0]), "padd")
        return u_big

    def unpadd(self, cc, cc_big):
        self._ck(self.lib.unpadd(self._ctx, _addr(cc, writable=True), _addr(cc_big), cc.shape[0]), "unpadd")
        return cc

    # -- module derivatives ---------------------------------------------------------------
    def ddx(self, f, dfdx):
        s = self.dims.shape
        self._ck(self.lib.ddx(self._ctx, _addr(f, s), _addr(dfdx, s, True)), "ddx")
        return dfdx

    def ddy(self, f, dfdy):
        s = self.dims.shape
        self._ck(self.lib.ddy(self._ctx, _addr(f, s), _addr(dfdy, s, True)), "ddy")
        return dfdy

    def ddxy(self, f, dfdx, dfdy):
        s = self.dims.shape
        self._ck(self.lib.ddxy(self._ctx, _addr(f, s), _addr(dfdx, s, True), _addr(dfdy, s, True)), "ddxy")
        return dfdx, dfdy

    def filt_da(self, f, dfdx, dfdy):
        """f is intent(inout): replaced by its Nyquist-filtered self (derivatives.f90:180)."""
        s = self.dims.shape
        self._ck(self.lib.filt_da(self._ctx, _addr(f, s, True), _addr(dfdx, s, True), _addr(dfdy, s, True)), "filt_da")
        return f, dfdx, dfdy

    def ddz_uv(self, f, dfdz):
        s = self.dims.shape
        self._ck(self.lib.ddz_uv(self._ctx, _addr(f, s), _addr(dfdz, s, True)), "ddz_uv")
        return dfdz

    def ddz_w(self, f, dfdz):
        s = self.dims.shape
        self._ck(self.lib.ddz_w(self._ctx, _addr(f, s), _addr(dfdz, s, True)), "ddz_w")
        return dfdz

    def test_filter(self, f, G):
        self._ck(self.lib.test_filter(self._ctx, _addr(f, writable=True), _addr(G), f.shape[0]), "test_filter")
        return f

    # -- convec / pressure ------------------------------------------------------------------
    def convec(self, u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy, RHSx, RHSy, RHSz):
        s = self.dims.shape
        ins = [_addr(a, s) for a in (u, v, w, dudy, dudz, dvdx, dvdz, dwdx, dwdy)]
        outs = [_addr(a, s, True) for a in (RHSx, RHSy, RHSz)]
        self._ck(self.lib.convec(self._ctx, *ins, *outs), "convec")
        return RHSx, RHSy, RHSz

    def press_stag_array(self, u, v, w, divtz, dt, tadv1, p, dpdx, dpdy, dpdz):
        s = self.dims.shape
        self._ck(self.lib.press_stag_array(self._ctx, _addr(u, s), _addr(v, s), _addr(w, s), _addr(divtz, s),
                                           float(dt), float(tadv1), _addr(p, s, True), _addr(dpdx, s, True),
                                           _addr(dpdy, s, True), _addr(dpdz, s, True)), "press_stag_array")
        return p, dpdx, dpdy, dpdz

    def tridag_array(self, a, b, c, r, u):
        n = r.shape[0]
        self._ck(self.lib.tridag_array(self._ctx, _addr(a), _addr(b), _addr(c), _addr(r), _addr(u, writable=True), n),
                 "tridag_array")
        return u

    # -- device-resident state ------------------------------------------------------------------
    def upload(self, name, host):
        self._ck(self.lib.upload(self._ctx, FIELD_IDS[name], _addr(host, self.dims.shape)), "upload")

    def download(self, name, out=None):
        out = self.empty() if out is None else out
        self._ck(self.lib.download(self._ctx, FIELD_IDS[name], _addr(out, self.dims.shape, True)), "download")
        return out

    def field_ptr(self, name) -> int:
        p = self.lib.field_ptr(self._ctx, FIELD_IDS[name])
        if not p:
            raise LibraryError("field_ptr: " + self.lib.error(self._ctx))
        return int(p)

    def step(self, dt, tadv1=1.5, tadv2=-0.5, first_step=False, mode=0, mean_p_force_x=0.0,
             mean_p_force_y=0.0, ubot=0.0, utop=0.0, nu=0.0, sgs_model=1, ifilter=1, Co=0.16,
             wall_damp_exp=2.0, vonk=0.4, zo=1e-4, lasd_cs_init=False, lasd_update=False, lasd_init_F=False,
             lagran_dt=0.0, turbines=False, turbines_eps=1.0):
        """One timestep main.f90:155-344 on the resident fields.  mode 0: core path (divt* as
        resident); mode 1: full step with wallstress, sgs_stag (constant coefficient, or sgs_model 5 =
        Lagrangian scale-dependent: lasd_* select the branch of sgs_stag_util.f90:183-216) and divstress."""
        sp = StepParams(dt, tadv1, tadv2, mean_p_force_x, mean_p_force_y, ubot, utop, nu, int(first_step), int(mode),
                        int(sgs_model), int(ifilter), Co, wall_damp_exp, vonk, zo, int(lasd_cs_init),
                        int(lasd_update), int(lasd_init_F), float(lagran_dt), int(turbines), float(turbines_eps))
        self._ck(self.lib.step(self._ctx, C.byref(sp)), "step")

    # -- restart file (io.f90:1173-1211, initial.f90:226-239) ------------------------------------------
    def checkpoint_write(self, fname):
        self._ck(self.lib.checkpoint_write(self._ctx, str(fname).encode()), "checkpoint_write")

    def checkpoint_read(self, fname):
        self._ck(self.lib.checkpoint_read(self._ctx, str(fname).encode()), "checkpoint_read")

    # -- running time averages (time_average.f90:176-320) --------------------------------------------
    def tavg_compute(self, dt):
        self._ck(self.lib.tavg_compute(self._ctx, float(dt)), "tavg_compute")

    def tavg_download(self, name):
        """Accumulator `name` (TAVG_IDS) as the (0:nz, ny, nx) array of tavg_t, and the accumulated time."""
        d = self.dims
        out = np.zeros((d.nz + 1, d.ny, d.nx))
        tt = C.c_double()
        self._ck(self.lib.tavg_download(self._ctx, TAVG_IDS[name], out.ctypes.data, C.byref(tt)), "tavg_download")
        return out, tt.value

    def tavg_reset(self):
        self._ck(self.lib.tavg_reset(self._ctx), "tavg_reset")

    # -- actuator disks (turbines.f90) --------------------------------------------------------------
    def turbines_init(self, farm, adm_correction=False, use_rotation=False, tip_speed_ratio=7.0):
        """farm: objects with nodes (n, 3) int (1-based i, j, local k), ind (n), nhat, Ct_prime, dia, M, u_d_T --
        what turbines_nodes (turbines.f90:275-462) leaves in wind_farm%turbine(:).  use_rotation (turbines.f90:76,
        :607-615): the objects also carry ind_t (n) and e_theta (n, 3)."""
        arr = (TurbineStruct * max(len(farm), 1))()
        keep = []
        for s, t in enumerate(farm):
            nodes = np.ascontiguousarray(np.asarray(t.nodes, dtype=np.int32).reshape(-1, 3))
            ind = np.ascontiguousarray(np.asarray(t.ind, dtype=np.float64))
            keep += [nodes, ind]
            arr[s].num_nodes = len(ind)
            arr[s].nodes = nodes.ctypes.data
            arr[s].ind = ind.ctypes.data
            arr[s].nhat = (C.c_double * 3)(*[float(x) for x in t.nhat])
            arr[s].Ct_prime, arr[s].dia, arr[s].M, arr[s].u_d_T = float(t.Ct_prime), float(t.dia), float(t.M), float(t.u_d_T)
        self._nturb = len(farm)
        self._ck(self.lib.turbines_init(self._ctx, len(farm), arr, int(adm_correction)), "turbines_init")
        if use_rotation:
            n = max(len(farm), 1)
            it, et = (C.c_void_p * n)(), (C.c_void_p * n)()
            for s, t in enumerate(farm):
                a = np.ascontiguousarray(np.asarray(t.ind_t, dtype=np.float64))
                b = np.ascontiguousarray(np.asarray(t.e_theta, dtype=np.float64).reshape(-1, 3))
                if len(a) != arr[s].num_nodes or len(b) != arr[s].num_nodes:
                    raise ValueError("ind_t / e_theta must have one entry per node")
                keep += [a, b]
                it[s], et[s] = a.ctypes.data, b.ctypes.data
            self._ck(self.lib.turbines_rotation(self._ctx, len(farm), it, et, float(tip_speed_ratio)), "turbines_rotation")

    def turbines_forcing(self, eps, fetch=True):
        """turbines_forcing (turbines.f90:465-638) on the resident fields; returns (u_d, u_d_T, f_n) per disk."""
        n = getattr(self, "_nturb", 0)
        if not fetch:
            self._ck(self.lib.turbines_forcing(self._ctx, float(eps), None, None, None), "turbines_forcing")
            return None
        out = [np.zeros(max(n, 1)) for _ in range(3)]
        self._ck(self.lib.turbines_forcing(self._ctx, float(eps), *[o.ctypes.data for o in out]), "turbines_forcing")
        return tuple(o[:n] for o in out)

    def max_cfl(self, dt):
        v = C.c_double()
        self._ck(self.lib.max_cfl(self._ctx, float(dt), C.byref(v)), "max_cfl")
        return v.value

    def cfl_dt(self, cfl):
        v = C.c_double()
        self._ck(self.lib.cfl_dt(self._ctx, float(cfl), C.byref(v)), "cfl_dt")
        return v.value

    def rmsdiv(self):
        v = C.c_double()
        self._ck(self.lib.rmsdiv(self._ctx, C.byref(v)), "rmsdiv")
        return v.value

    def host_register(self, arr):
        """Page-lock a numpy array that will be passed to the per-routine entry points repeatedly."""
        self._ck(self.lib.host_register(self._ctx, arr.ctypes.data, arr.nbytes), "host_register")

    def host_unregister(self, arr):
        self._ck(self.lib.host_unregister(self._ctx, arr.ctypes.data), "host_unregister")

    # -- multi-GPU ---------------------------------------------------------------------------------
    def comm_unique_id(self, local: bool = False) -> bytes:
        """128-byte id for comm_init: NCCL's (one rank per GPU), or with local=True the id of the
        single-device transport (all ranks are threads of this process on ONE GPU)."""
        buf = C.create_string_buffer(128)
        if (self.lib.comm_local_id if local else self.lib.comm_unique_id)(buf):
            raise LibraryError("comm_unique_id: " + self.lib.error(None))
        return buf.raw

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.comm_init(self._ctx, buf), "comm_init")

    def comm_p2p_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.comm_p2p_export(self._ctx, buf), "comm_p2p_export")
        return buf.raw

    def comm_p2p_import(self, blobs):
        """blobs: the nproc 128-byte exports in rank order; switches the pressure transposes to peer memory."""
        if blobs is None:                       # back to the NCCL all-to-alls
            self._ck(self.lib.comm_p2p_import(self._ctx, None), "comm_p2p_import")
            self.p2p_enabled = False
            return
        raw = b"".join(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        self._ck(self.lib.comm_p2p_import(self._ctx, buf), "comm_p2p_import")
        self.p2p_enabled = True

    def sync_real_array(self, var, isync=3):
        self._ck(self.lib.sync_real_array(self._ctx, _addr(var, self.dims.shape, True), int(isync)), "sync_real_array")
        return var
