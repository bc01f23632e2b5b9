"""The ISO_C_BINDING shims in fortran/ cannot be compiled in this pool (no Fortran compiler on any box), so
their agreement with the C ABI is checked textually: every bind(C) interface must name a function include/lesgo_gpu.h
declares, with the same number of arguments, every bind(C) derived type must list the members of the C struct of the same
name in the same order with matching kinds, and every lesgo_gpu_* call in the shims must go through a declared interface."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lesgo_gpu.h")
FORTRAN = os.path.join(ROOT, "fortran")


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def c_prototypes():
    """name -> number of parameters, for every function include/lesgo_gpu.h declares."""
    text = _strip_c(open(HEADER).read())
    out = {}
    for m in re.finditer(r"\b(lesgo_gpu_\w+|dfftw_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def c_structs():
    """struct name -> [(kind, member), ...] with kind in {int, double, ptr}."""
    text = _strip_c(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*\w+\s*;", text, flags=re.S):
        members = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            kind = "ptr" if "*" in decl else ("double" if re.match(r"(const\s+)?double\b", decl) else "int")
            names = re.sub(r"^(const\s+)?(int|double|long|unsigned|void)\b[\s\*]*(const\s+)?", "", decl)
            for n in names.split(","):
                n = n.strip().lstrip("*").strip()
                arr = re.match(r"(\w+)\s*\[(\d+)\]", n)
                members.append((kind, arr.group(1) if arr else n, int(arr.group(2)) if arr else 1))
        out[m.group(1)] = members
    return out


def fortran_logical_lines(path):
    lines, cur = [], ""
    for raw in open(path):
        ln = raw.split("!")[0].rstrip() if "'" not in raw.split("!")[0] or raw.count("'") % 2 == 0 else raw.rstrip()
        if not ln.strip():
            continue
        ln = ln.strip()
        if ln.startswith("&"):
            ln = ln[1:].lstrip()
        if ln.endswith("&"):
            cur += ln[:-1] + " "
            continue
        lines.append(cur + ln)
        cur = ""
    return lines


def fortran_files():
    return sorted(os.path.join(FORTRAN, f) for f in os.listdir(FORTRAN) if f.endswith(".f90"))


def fortran_interfaces():
    """Fortran interface name -> (file, number of dummy arguments, bind(C) name); two interfaces may bind one C symbol
    (lesgo_gpu_comm_p2p_import and its ..._null variant that passes a null pointer)."""
    out = {}
    for path in fortran_files():
        for ln in fortran_logical_lines(path):
            m = re.search(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)", ln, flags=re.I)
            if m:
                args = [a for a in m.group(2).split(",") if a.strip()]
                out[m.group(1)] = (os.path.basename(path), len(args), m.group(3))
    return out


def fortran_types():
    out = {}
    for path in fortran_files():
        lines = fortran_logical_lines(path)
        i = 0
        while i < len(lines):
            m = re.match(r"type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*(\w+)", lines[i], flags=re.I)
            if m:
                members = []
                i += 1
                while not re.match(r"end\s+type", lines[i], flags=re.I):
                    d = re.match(r"(integer\s*\(\s*c_int\s*\)|real\s*\(\s*c_double\s*\)|type\s*\(\s*c_ptr\s*\))\s*::\s*(.*)", lines[i], flags=re.I)
                    assert d, (path, lines[i])
                    kind = {"i": "int", "r": "double", "t": "ptr"}[d.group(1)[0].lower()]
                    for n in re.findall(r"(\w+)\s*(?:\(\s*(\d+)\s*\))?\s*(?:,|$)", d.group(2)):
                        members.append((kind, n[0], int(n[1]) if n[1] else 1))
                    i += 1
                out[m.group(1)] = members
            i += 1
    return out


def test_every_fortran_interface_matches_a_c_prototype():
    protos, ifaces = c_prototypes(), fortran_interfaces()
    assert len(protos) >= 50 and len(ifaces) >= 35, (len(protos), len(ifaces))
    for fname_, (fname, nargs, name) in ifaces.items():
        assert name in protos, f"{fname}: bind(C) name {name} is not declared in include/lesgo_gpu.h"
        assert nargs == protos[name], f"{fname}: {fname_} has {nargs} dummy arguments, the C prototype of {name} {protos[name]}"


def test_fortran_bind_c_types_mirror_the_c_structs():
    cs, fs = c_structs(), fortran_types()
    assert {"lesgo_gpu_dims", "lesgo_gpu_step_params", "lesgo_gpu_turbine"} <= set(fs), sorted(fs)
    for name, members in fs.items():
        assert name in cs, name
        got = [(k, n.lower(), c) for k, n, c in members]
        want = [(k, n.lower(), c) for k, n, c in cs[name]]
        assert got == want, (name, got, want)


def test_every_call_in_the_shims_has_an_interface():
    ifaces = {n.lower() for n in fortran_interfaces()}
    for path in fortran_files():
        text = "\n".join(fortran_logical_lines(path))
        for m in re.finditer(r"\b(lesgo_gpu_\w+)\s*\(", text):
            name = m.group(1).lower()
            if name in ("lesgo_gpu_dims", "lesgo_gpu_step_params", "lesgo_gpu_turbine"):
                continue
            assert name in ifaces, f"{os.path.basename(path)}: {name} is called but has no bind(C) interface"
