"""ctypes binding of the C ABI declared in include/lesgo_gpu.h."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class LibraryError(RuntimeError):
    pass


class DimsStruct(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nz_tot", C.c_int),
                ("nproc", C.c_int), ("coord", C.c_int),
                ("L_x", C.c_double), ("L_y", C.c_double), ("dz", C.c_double),
                ("lbc_mom", C.c_int), ("ubc_mom", C.c_int), ("sgs", C.c_int), ("device", C.c_int)]


class StepParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("tadv1", C.c_double), ("tadv2", C.c_double),
                ("mean_p_force_x", C.c_double), ("mean_p_force_y", C.c_double),
                ("ubot", C.c_double), ("utop", C.c_double), ("nu_molec_nd", C.c_double),
                ("first_step", C.c_int), ("mode", C.c_int),
                ("sgs_model", C.c_int), ("ifilter", C.c_int),
                ("Co", C.c_double), ("wall_damp_exp", C.c_double), ("vonk", C.c_double), ("zo", C.c_double),
                ("lasd_cs_init", C.c_int), ("lasd_update", C.c_int), ("lasd_init_F", C.c_int),
                ("lagran_dt", C.c_double), ("turbines", C.c_int), ("turbines_eps", C.c_double)]


class TurbineStruct(C.Structure):
    _fields_ = [("num_nodes", C.c_int), ("nodes", C.c_void_p), ("ind", C.c_void_p), ("nhat", C.c_double * 3),
                ("Ct_prime", C.c_double), ("dia", C.c_double), ("M", C.c_double), ("u_d_T", C.c_double)]


# every symbol include/lesgo_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.c_void_p          # double* (host or device address)
SYMBOLS = {
    "lesgo_gpu_create": (C.c_int, [C.POINTER(DimsStruct), C.POINTER(_P)]),
    "lesgo_gpu_destroy": (C.c_int, [_P]),
    "lesgo_gpu_last_error": (C.c_char_p, [_P]),
    "lesgo_gpu_set_stream": (C.c_int, [_P, _P]),
    "lesgo_gpu_host_register": (C.c_int, [_P, _P, C.c_size_t]),
    "lesgo_gpu_host_unregister": (C.c_int, [_P, _P]),
    "lesgo_gpu_synchronize": (C.c_int, [_P]),
    "lesgo_gpu_launch_count": (C.c_long, [_P]),
    "lesgo_gpu_profile": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int]),
    "lesgo_gpu_wavenumbers": (C.c_int, [_P, _D, _D, _D]),
    "lesgo_gpu_padd": (C.c_int, [_P, _D, _D, C.c_int]),
    "lesgo_gpu_unpadd": (C.c_int, [_P, _D, _D, C.c_int]),
    "lesgo_gpu_fft_r2c": (C.c_int, [_P, _D, _D, C.c_int, C.c_int]),
    "lesgo_gpu_fft_c2r": (C.c_int, [_P, _D, _D, C.c_int, C.c_int]),
    "lesgo_gpu_fftw_bind": (C.c_int, [_P, C.POINTER(DimsStruct)]),
    "lesgo_gpu_fftw_plan_2d": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "lesgo_gpu_fftw_execute": (C.c_int, [C.c_longlong, C.c_int, _D, _D]),
    "lesgo_gpu_fftw_destroy": (C.c_int, [C.c_longlong]),
    "lesgo_gpu_fftw_last_error": (C.c_char_p, []),
    "lesgo_gpu_ddx": (C.c_int, [_P, _D, _D]),
    "lesgo_gpu_ddy": (C.c_int, [_P, _D, _D]),
    "lesgo_gpu_ddxy": (C.c_int, [_P, _D, _D, _D]),
    "lesgo_gpu_filt_da": (C.c_int, [_P, _D, _D, _D]),
    "lesgo_gpu_ddz_uv": (C.c_int, [_P, _D, _D]),
    "lesgo_gpu_ddz_w": (C.c_int, [_P, _D, _D]),
    "lesgo_gpu_test_filter": (C.c_int, [_P, _D, _D, C.c_int]),
    "lesgo_gpu_convec": (C.c_int, [_P] + [_D] * 12),
    "lesgo_gpu_press_stag_array": (C.c_int, [_P, _D, _D, _D, _D, C.c_double, C.c_double, _D, _D, _D, _D]),
    "lesgo_gpu_tridag_array": (C.c_int, [_P, _D, _D, _D, _D, _D, C.c_int]),
    "lesgo_gpu_field_ptr": (_P, [_P, C.c_int]),
    "lesgo_gpu_upload": (C.c_int, [_P, C.c_int, _D]),
    "lesgo_gpu_download": (C.c_int, [_P, C.c_int, _D]),
    "lesgo_gpu_step": (C.c_int, [_P, C.POINTER(StepParams)]),
    "lesgo_gpu_max_cfl": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double)]),
    "lesgo_gpu_cfl_dt": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double)]),
    "lesgo_gpu_rmsdiv": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "lesgo_gpu_checkpoint_write": (C.c_int, [_P, C.c_char_p]),
    "lesgo_gpu_checkpoint_read": (C.c_int, [_P, C.c_char_p]),
    "lesgo_gpu_tavg_compute": (C.c_int, [_P, C.c_double]),
    "lesgo_gpu_tavg_download": (C.c_int, [_P, C.c_int, _D, C.POINTER(C.c_double)]),
    "lesgo_gpu_tavg_reset": (C.c_int, [_P]),
    "lesgo_gpu_turbines_init": (C.c_int, [_P, C.c_int, C.POINTER(TurbineStruct), C.c_int]),
    "lesgo_gpu_turbines_rotation": (C.c_int, [_P, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double]),
    "lesgo_gpu_turbines_forcing": (C.c_int, [_P, C.c_double, _D, _D, _D]),
    "lesgo_gpu_comm_unique_id": (C.c_int, [_P]),
    "lesgo_gpu_comm_local_id": (C.c_int, [_P]),
    "lesgo_gpu_comm_init": (C.c_int, [_P, _P]),
    "lesgo_gpu_comm_p2p_export": (C.c_int, [_P, _P]),
    "lesgo_gpu_comm_p2p_import": (C.c_int, [_P, _P]),
    "lesgo_gpu_sync_real_array": (C.c_int, [_P, _D, C.c_int]),
}


# FFTW3 legacy-Fortran symbols the library exports for the reference's remaining CPU callers (by reference)
_LL = C.POINTER(C.c_longlong)
_I = C.POINTER(C.c_int)
FFTW_SYMBOLS = {
    "dfftw_plan_dft_r2c_2d_": (None, [_LL, _I, _I, _D, _D, _I]),
    "dfftw_plan_dft_c2r_2d_": (None, [_LL, _I, _I, _D, _D, _I]),
    "dfftw_execute_dft_r2c_": (None, [_LL, _D, _D]),
    "dfftw_execute_dft_c2r_": (None, [_LL, _D, _D]),
    "dfftw_destroy_plan_": (None, [_LL]),
}


def library_path() -> str:
    # LESGO_CUDA_LIB: developer override to A/B an experimental build of the same CUDA library
    return os.environ.get("LESGO_CUDA_LIB") or os.path.join(_HERE, "liblesgo_cuda.so")


class Library:
    """A loaded liblesgo_cuda.so with typed entry points."""

    def __init__(self, path: str | None = None):
        self.path = path or library_path()
        if not os.path.exists(self.path):
            raise LibraryError(
                f"{self.path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  lesgo_b200 has no CPU fallback.")
        try:
            self.dll = C.CDLL(self.path)
        except OSError as e:
            raise LibraryError(f"cannot load {self.path}: {e}") from e
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError as e:
                raise LibraryError(f"{self.path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("lesgo_gpu_"):], fn)
        for name, (res, args) in FFTW_SYMBOLS.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError as e:
                raise LibraryError(f"{self.path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
            setattr(self, name.rstrip("_"), fn)

    def error(self, ctx=None) -> str:
        s = self.last_error(ctx)
        return s.decode() if s else ""


_LIB = None


def load_library() -> Library:
    global _LIB
    if _LIB is None:
        _LIB = Library()
    return _LIB
