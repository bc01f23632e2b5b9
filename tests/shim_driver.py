"""The reference's OWN time loop (main.f90, interpreted by oracle/f90exec.py) driving the ISO_C_BINDING shims of fortran/
in place of the five sources they replace, with every `bind(C)` call carried through ctypes into a liblesgo_cuda build
(the kernel-logic emulator on CPU, the sm_100a library on a B200).  No box of this pool has a Fortran compiler, so this is
how the drop-in boundary gets executed at all: derivatives.f90, convec.f90, press_stag_array.f90, tridag_array.f90 and
fft.f90 of fortran/ are parsed and run statement by statement, their arguments marshalled exactly as gfortran would pass
them (base address of a contiguous array, scalars by value where the interface says `value`).

What is NOT exercised: gpu_require (context creation, MPI / NCCL bootstrap) -- the context is made here and handed in --
and gpu_check's error-string plumbing; both are replaced by Python stand-ins.  Test infrastructure, like oracle/."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import f90exec as F  # noqa: E402
from oracle import refrun  # noqa: E402
import lesgo_b200  # noqa: E402
from lesgo_b200 import lib as L  # noqa: E402

FORTRAN = os.path.join(ROOT, "fortran")
REPLACED = ("fft.f90", "derivatives.f90", "convec.f90", "tridag_array.f90", "press_stag_array.f90")

ISO_C_BINDING = """module iso_c_binding
implicit none
integer, parameter :: c_int = 4, c_double = 8, c_intptr_t = 8, c_char = 1, c_long = 8, c_size_t = 8
type :: c_ptr
    integer :: addr
end type c_ptr
end module iso_c_binding
module mpi
implicit none
integer, parameter :: MPI_COMM_TYPE_SHARED = 1, MPI_INFO_NULL = 0, MPI_CHARACTER = 1, MPI_INTEGER = 2, MPI_MIN = 3
end module mpi
"""


class CPtr:
    """type(c_ptr): an address."""

    def __init__(self, addr=0):
        self.addr = int(addr)

    def __repr__(self):
        return f"CPtr({self.addr:#x})"


_KEEP = []          # int32 images of the interpreter's integer arrays (it keeps them as int64), alive while C reads them


def _c_image(a):
    """The memory C sees for a Fortran array: real(c_double) arrays as they are, integer(c_int) arrays as an int32 copy."""
    assert a.flags.f_contiguous, "the shims pass whole contiguous arrays"
    if a.dtype == np.float64 or a.dtype == np.uint8:
        return a
    assert a.dtype.kind == "i", a.dtype
    img = np.asfortranarray(a.astype(np.int32))
    _KEEP.append(img)
    return img


def _addr_of(x):
    if isinstance(x, CPtr):
        return x.addr
    if isinstance(x, F.FArray):
        return _c_image(x.a).ctypes.data
    if isinstance(x, F.ElemRef):
        img = _c_image(x.base.a)
        return img.ctypes.data + int(np.ravel_multi_index(x.idx, img.shape, order="F")) * img.itemsize
    raise TypeError(f"cannot take the address of {type(x).__name__}")


def _vals(a):
    """Arguments as values: `call x(...)` hands externals (value, setter) pairs, a function reference plain values."""
    return [x[0] if isinstance(x, tuple) else x for x in a]


class ShimmedReference(refrun.Reference):
    """refrun.Reference with the five replaced sources taken from fortran/ and lesgo_gpu_* bound to `core`'s library."""

    def __init__(self, p, core=None, files=refrun.FILES, resident=False, lib=None, **kw):
        """resident: also load fortran/lesgo_gpu_resident_mod.f90 (whole-step entry points, LASD switches).
        core = None (with lib = a loaded Library): nothing is pre-made -- the shims' own gpu_require creates the context
        through lesgo_gpu_create, gpu_pin registers the arrays, gpu_check is the Fortran one."""
        self.core = core
        self.lib = lib if lib is not None else core.lib
        self.ctx = None
        self.calls = {}
        if resident:
            files = list(files) + [os.path.join(FORTRAN, "lesgo_gpu_resident_mod.f90")]
        tmp = tempfile.NamedTemporaryFile("w", suffix=".f90", delete=False)
        tmp.write(ISO_C_BINDING)
        tmp.close()
        self._iso = tmp.name
        order = []
        for f in files:
            if f == "fft.f90":                    # lesgo_gpu_mod needs param + messages, the shims need lesgo_gpu_mod
                order += [self._iso, os.path.join(FORTRAN, "lesgo_gpu_mod.f90")]
            order.append(f)
        overrides = {f: os.path.join(FORTRAN, f) for f in REPLACED}
        try:
            super().__init__(p, files=order, overrides=overrides, post_load=ShimmedReference._bind, **kw)
        finally:
            os.unlink(self._iso)

    # ---- the C side of every bind(C) interface ------------------------------------------------------------------
    def _bind(self):
        I, core, lib = self.I, self.core, self.lib
        mod = I.modules["lesgo_gpu_mod"]
        I.set("iso_c_binding", "c_null_ptr", CPtr(0))
        I.set("iso_c_binding", "c_null_char", "\x00")
        ext = I.externals
        if core is not None:
            for name in ("gpu_require", "gpu_check", "gpu_pin", "gpu_pin_sim_param"):
                mod.procs.pop(name, None)         # Python stand-ins below (externals are consulted after module procs)
            ctx = CPtr(core._ctx.value if hasattr(core._ctx, "value") else int(core._ctx))
            I.set("lesgo_gpu_mod", "gpu_ctx", ctx)
            core._ctx_ptr = lambda: ctx
            self.ctx = ctx

            def gpu_check(fr, a):
                rc, where = _vals(a)[:2]
                if int(rc) != 0:
                    raise lesgo_b200.LibraryError(f"{where}: {lib.error(C.c_void_p(ctx.addr))}")

            ext["gpu_require"] = lambda fr, a: None
            ext["gpu_pin"] = lambda fr, a: None
            ext["gpu_pin_sim_param"] = lambda fr, a: None
            ext["gpu_check"] = gpu_check
        else:
            # the shims' own start-up path: gpu_require -> lesgo_gpu_create(lesgo_gpu_dims(...), gpu_ctx)
            I.set("lesgo_gpu_mod", "gpu_ctx", CPtr(0))
            td = mod.types["lesgo_gpu_dims"]

            def dims_constructor(fr, a):
                o = F.FStruct(td)
                for name, v in zip(td.members, _vals(a)):
                    setattr(o, name, v)
                return o

            def create(fr, a):
                d = _vals(a)[0]
                st = L.DimsStruct()
                for fname, ftype in L.DimsStruct._fields_:
                    setattr(st, fname, float(getattr(d, fname.lower())) if ftype is C.c_double else int(getattr(d, fname.lower())))
                out = C.c_void_p()
                rc = int(lib.create(C.byref(st), C.byref(out)))
                self.ctx = CPtr(out.value or 0)
                I.set("lesgo_gpu_mod", "gpu_ctx", self.ctx)
                self.calls["lesgo_gpu_create"] = self.calls.get("lesgo_gpu_create", 0) + 1
                return rc

            def out_int(value):
                def f(fr, a):
                    a[-2][1](value)               # (..., result, ierr): the result is the last argument but one
                return f
            ext["lesgo_gpu_dims"] = dims_constructor
            ext["mpi_comm_split_type"] = out_int(1)
            ext["mpi_comm_rank"] = out_int(self.p.coord)      # all ranks of a test share one node: local rank = coord
            ext["mpi_comm_free"] = lambda fr, a: None
            # collectives gpu_require uses on byte buffers and flags, over the mailboxes of refrun.run_ranks
            import queue
            me, nproc, boxes = self.p.coord, self.p.nproc, self.boxes

            def box(src, dst, tag):
                with boxes["lock"]:
                    return boxes.setdefault((src, dst, tag), queue.Queue())

            def bcast(fr, a):
                buf, count, root = a[0][0], int(a[1][0]), int(a[3][0])
                flat = buf.a.reshape(-1, order="F")
                if me == root:
                    for r in range(nproc):
                        if r != me:
                            box(me, r, "bcast").put(np.array(flat[:count], copy=True))
                else:
                    flat[:count] = box(root, me, "bcast").get(timeout=refrun.MPI_TIMEOUT)

            def allgather(fr, a):
                send, scount, recv = a[0][0], int(a[1][0]), a[3][0]
                mine = np.array(send.a.reshape(-1, order="F")[:scount], copy=True)
                for r in range(nproc):
                    if r != me:
                        box(me, r, "allgather").put(mine)
                out = recv.a.reshape(-1, order="F")
                for r in range(nproc):
                    out[r * scount:(r + 1) * scount] = mine if r == me else box(r, me, "allgather").get(timeout=refrun.MPI_TIMEOUT)

            def allreduce_min_int(fr, a):
                # gpu_require only reduces its 0 / 1 success flags with MPI_MIN
                mine = int(a[0][0])
                for r in range(nproc):
                    if r != me:
                        box(me, r, "flag").put(mine)
                acc = mine
                for r in range(nproc):
                    if r != me:
                        acc = min(acc, box(r, me, "flag").get(timeout=refrun.MPI_TIMEOUT))
                a[1][1](acc)
            if nproc > 1:
                ext["mpi_bcast"] = bcast
                ext["mpi_allgather"] = allgather
                self._flag_allreduce = allreduce_min_int
        ext["c_associated"] = lambda fr, a: _vals(a)[0].addr != 0
        ext["c_loc"] = lambda fr, a: CPtr(_addr_of(_vals(a)[0]))

        def c_f_pointer(fr, a):
            # call c_f_pointer(cptr, fptr, shape) for a character(kind=c_char) pointer array: the C string, NUL padded
            n = int(np.asarray(getattr(a[2][0], "a", a[2][0])).reshape(-1)[0])
            raw = C.string_at(a[0][0].addr)[:n]
            chars = np.empty(n, dtype=object)
            for i in range(n):
                chars[i] = chr(raw[i]) if i < len(raw) else "\x00"
            a[1][1](F.FArray(chars, (1,), "object"))
        ext["c_f_pointer"] = c_f_pointer

        def last_error(fr, a):
            fn = lib.dll.lesgo_gpu_last_error
            fn.restype, fn.argtypes = C.c_void_p, [C.c_void_p]
            return CPtr(fn(C.c_void_p(_vals(a)[0].addr)) or 0)

        def transfer(fr, a):
            src, mold = _vals(a)[:2]
            if isinstance(mold, CPtr):
                return CPtr(int(src))
            return _addr_of(src) if isinstance(src, (CPtr, F.FArray, F.ElemRef)) else src
        ext["transfer"] = transfer

        def c_function(cname):
            restype, argtypes = L.SYMBOLS[cname]
            fn = getattr(lib.dll, cname)
            fn.restype, fn.argtypes = restype, argtypes

            def call(fr, a):
                vals = _vals(a)
                assert len(vals) == len(argtypes), (cname, len(vals), len(argtypes))
                cargs = []
                for v, t in zip(vals, argtypes):
                    if isinstance(v, F.FArray) and v.a.dtype == object:      # an array of bind(C) structs or of c_ptr
                        elems = list(v.a.reshape(-1, order="F"))
                        if all(isinstance(e, CPtr) for e in elems):
                            arr = (C.c_void_p * max(len(elems), 1))(*[e.addr for e in elems])
                        else:
                            arr = (L.TurbineStruct * max(len(elems), 1))()
                            for i, e in enumerate(elems):
                                arr[i].num_nodes = int(e.num_nodes)
                                arr[i].nodes, arr[i].ind = e.nodes.addr, e.ind.addr
                                arr[i].nhat = (C.c_double * 3)(*[float(z) for z in np.asarray(getattr(e.nhat, "a", e.nhat)).reshape(-1)])
                                arr[i].Ct_prime, arr[i].dia, arr[i].M, arr[i].u_d_T = float(e.ct_prime), float(e.dia), float(e.m), float(e.u_d_t)
                        _KEEP.append(arr)
                        cargs.append(arr if t is not C.c_void_p else C.cast(arr, C.c_void_p))   # ctypes converts the array
                        continue
                    if isinstance(v, str):           # trim(name) // c_null_char: a NUL-terminated C string
                        buf = C.create_string_buffer(v.split("\x00")[0].encode())
                        _KEEP.append(buf)
                        cargs.append(C.cast(buf, C.c_void_p) if t is C.c_void_p else buf.value)
                        continue
                    if isinstance(v, F.FStruct):     # a bind(C) derived type, passed by reference: member by member
                        cls = {"lesgo_gpu_step_params": L.StepParams, "lesgo_gpu_dims": L.DimsStruct}[v._type.name.lower()]
                        st = cls()
                        for fname, ftype in cls._fields_:
                            val = getattr(v, fname.lower())
                            setattr(st, fname, (float(val) if ftype is C.c_double else int(val)) if val is not None else 0)
                        cargs.append(C.byref(st))
                        continue
                    if t in (C.c_int, C.c_long, C.c_longlong, C.c_size_t, C.c_ulong, C.c_ulonglong, C.c_uint):
                        cargs.append(int(v))
                    elif t is C.c_double:
                        cargs.append(float(v))
                    else:                           # pointers: the context, or the base address of an array
                        cargs.append(C.c_void_p(_addr_of(v)))
                self.calls[cname] = self.calls.get(cname, 0) + 1
                return int(fn(*cargs))
            return call

        for cname in L.SYMBOLS:
            ext[cname] = c_function(cname)
        ext["lesgo_gpu_last_error"] = last_error
        if core is None:
            ext["lesgo_gpu_create"] = create
            if self.p.nproc > 1:
                base_allreduce = ext["mpi_allreduce"]

                def allreduce(fr, a):
                    if not isinstance(a[0][0], F.FArray) and a[4][0] == 3:      # MPI_MIN of module mpi above
                        return self._flag_allreduce(fr, a)
                    return base_allreduce(fr, a)
                ext["mpi_allreduce"] = allreduce
        ext["press_raw"] = ext["lesgo_gpu_press_stag_array"]      # the local interface name of fortran/press_stag_array.f90
