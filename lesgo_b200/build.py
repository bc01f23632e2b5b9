"""In-tree build of liblesgo_cuda.so: nvcc, sm_100a only, one object per .cu in parallel."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
TARGET = os.path.join(HERE, "liblesgo_cuda.so")
SOURCES = ["lesgo_gpu.cu", "comm.cu", "xfwd_scale.cu", "xfwd_vort.cu", "xfwd_convec.cu", "xinv.cu", "ypass.cu", "prodfwd.cu", "fftw_shim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newest_header():
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith(".h"):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    inc = os.path.join(os.path.dirname(HERE), "include", "lesgo_gpu.h")
    return max(t, os.path.getmtime(inc))


def _compile(src, log_dir, defines=()):
    obj = os.path.join(log_dir, src.replace(".cu", ".o"))
    cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(log_dir, src + ".ptxas.log"), "w") as f:
        f.write(r.stdout)
    if r.returncode:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout[-4000:]}")
    return obj


def build(force: bool = False, verbose: bool = True, defines=(), target: str = TARGET, obj_dir: str = OBJ) -> str:
    """defines/target/obj_dir let a developer build an experimental variant next to the product library."""
    if shutil.which("nvcc") is None:
        raise RuntimeError("nvcc not found: liblesgo_cuda.so can only be built with the CUDA toolkit")
    OBJ = obj_dir
    TARGET = target
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _newest_header()
    todo = []
    for s in SOURCES:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        src_t = max(os.path.getmtime(os.path.join(CSRC, s)), hdr_t)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_t:
            todo.append(s)
    if todo:
        if verbose:
            print(f"[lesgo_b200.build] nvcc sm_100a: {', '.join(todo)}", file=sys.stderr)
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, OBJ, defines), todo))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if todo or not os.path.exists(TARGET):
        cmd = ["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-o", TARGET, *objs, "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError("link failed:\n" + r.stdout)
    return TARGET


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    if defs:
        tag = "_".join(d.lower() for d in defs)
        print(build(force="--force" in sys.argv, defines=defs, target=os.path.join(HERE, f"liblesgo_cuda.{tag}.so"),
                    obj_dir=os.path.join(CSRC, "_obj_" + tag)))
    else:
        print(build(force="--force" in sys.argv))
