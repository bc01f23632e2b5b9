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


def test_library_exports_every_header_symbol(lib):
    for name in header_symbols():
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
