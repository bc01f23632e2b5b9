"""The C-ABI library must exist in-tree, load, and export every symbol that
include/lesgo_gpu.h declares (no compute here: this runs without a GPU)."""
import os
import re
import shutil

import pytest

import lesgo_b200
from lesgo_b200.lib import SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lesgo_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lesgo_gpu_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(lesgo_b200.library_path()):
        if shutil.which("nvcc") is None:
            pytest.skip("liblesgo_cuda.so not built and no nvcc")
        from lesgo_b200 import build
        build.build()
    return lesgo_b200.Library()


def test_header_and_binding_agree():
    hs = header_symbols()
    assert hs, "no symbols parsed from the header"
    assert sorted(SYMBOLS) == hs


def header_fftw_symbols():
    txt = open(os.path.join(ROOT, "include", "lesgo_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfftw_[a-z0-9_]+_)\s*\(", txt)))


def test_library_exports_every_header_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib.dll, name), name
    from lesgo_b200.lib import FFTW_SYMBOLS
    fs = header_fftw_symbols()
    assert len(fs) == 5 and sorted(FFTW_SYMBOLS) == fs
    for name in fs:                      # the gfortran-mangled FFTW3 legacy API
        assert hasattr(lib.dll, name), name


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lesgo_b200.LibraryError, match="no CUDA device"):
        lesgo_b200.Core(lesgo_b200.Dims(nx=16, ny=16, Nz=4))


def test_product_package_does_not_touch_oracle_or_emulator():
    pkg = os.path.join(ROOT, "lesgo_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|lesgo_oracle", src, re.M), f
                assert "liblesgo_emul" not in src, f


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct of include/lesgo_gpu.h, taken from the C compiler, against the
    ctypes mirrors in lesgo_b200/lib.py (the Fortran bind(C) types in fortran/ list the same members)."""
    import ctypes as C
    import subprocess
    from lesgo_b200.lib import DimsStruct, StepParams, TurbineStruct
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"lesgo_gpu_dims": DimsStruct, "lesgo_gpu_step_params": StepParams, "lesgo_gpu_turbine": TurbineStruct}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lesgo_gpu.h"\nint main(void) {\n'
                   + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
