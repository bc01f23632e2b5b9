"""lesgo_b200 -- B200-native (sm_100a) implementation of LESGO's per-timestep
pseudo-spectral core behind the reference's own subroutine interface.

The compute path is `liblesgo_cuda.so` (hand-written FP64 CUDA, C ABI in
`include/lesgo_gpu.h`); this package is the thin Python host mirror of that interface
used by the tests and the benchmark.  There is no CPU fallback: importing works
anywhere, but creating a `Core` without the CUDA library or without a GPU raises.
"""
from .lib import Library, LibraryError, load_library, library_path  # noqa: F401
from .core import Core, Dims, FIELD_IDS, TAVG_IDS  # noqa: F401
from . import slab  # noqa: F401

__all__ = ["Library", "LibraryError", "load_library", "library_path", "Core", "Dims", "FIELD_IDS", "TAVG_IDS"]
