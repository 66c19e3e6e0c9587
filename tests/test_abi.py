"""CPU-only: the C-ABI library loads, exports every symbol include/ocean_b200.h declares, the ctypes struct layouts agree
with the C structs, and the product fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ocean_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"^(?:int32_t|const char \*)\s*(ob_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import ocean_b200 as ob
    lib = ob.lib()
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), "libocean_b200.so does not export %s" % n


def test_python_prototypes_cover_the_header():
    from ocean_b200 import _abi
    bound = set(_abi.PROTOTYPES) | set(_abi._STR)
    assert set(_declared()) == bound, (sorted(set(_declared()) - bound), sorted(bound - set(_declared())))


def test_struct_layouts_match_the_c_header():
    from ocean_b200 import _abi
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ocean_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ob_grid_desc), sizeof(ob_bc_desc), sizeof(ob_closure_desc), sizeof(ob_model_desc),
         offsetof(ob_model_desc, closures), offsetof(ob_model_desc, bcs_u), offsetof(ob_model_desc, bcs_kappae));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    M = _abi.ModelDesc
    want = [C.sizeof(_abi.GridDesc), C.sizeof(_abi.BcDesc), C.sizeof(_abi.ClosureDesc), C.sizeof(M),
            M.closures.offset, M.bcs_u.offset, M.bcs_kappae.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device():
    import ocean_b200 as ob
    n = C.c_int32(-1)
    status = ob.lib().ob_device_count(C.byref(n))
    if status == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ob.OceanB200Error):
        ob.B200()
    with pytest.raises(NotImplementedError):
        ob.CPU()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oceananigans.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in txt and "import oracle" not in txt, f


# ---------------------------------------------------------------------------------------------------------------
# The Julia shim cannot be executed here (no Julia): check it statically against the header instead.
# ---------------------------------------------------------------------------------------------------------------
JL = os.path.join(ROOT, "oceananigans.jl_b200", "julia", "OceananigansB200Ext.jl")
HDR = os.path.join(ROOT, "include", "ocean_b200.h")


def _c_structs():
    """{struct name: [(field name, C type, array dims)]} from include/ocean_b200.h"""
    import re
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(OB_MAX_\w+)\s+(\d+)", txt)}
    out = {}
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(\w+);", txt, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            if decl.startswith("const void *"):
                ctype, names = "const void *", decl[len("const void *"):]
            else:
                ctype, names = decl.split(" ", 1)
            for nm in names.split(","):
                nm = nm.strip()
                dims = [macros.get(d, None) if not d.isdigit() else int(d) for d in re.findall(r"\[(\w+)\]", nm)]
                fields.append((re.sub(r"\[.*", "", nm), ctype, dims))
        out[m.group(2)] = fields
    return out


def _jl_structs():
    """{struct name: [(field name, Julia type)]} from the shim"""
    import re
    txt = open(JL).read()
    out = {}
    for m in re.finditer(r"^struct (Ob\w+)\s*;?(.*?)\bend$", txt, flags=re.S | re.M):
        body = m.group(2).replace("\n", ";")
        fields = []
        for part in body.split(";"):
            part = part.strip()
            if "::" in part:
                n, t = part.split("::", 1)
                fields.append((n.strip(), t.strip()))
        out[m.group(1)] = fields
    return out


def _jl_type(ctype, dims, structs):
    base = {"int32_t": "Int32", "double": "Float64", "const void *": "Ptr{Cvoid}"}.get(ctype) or structs[ctype]
    for d in reversed(dims):
        base = "NTuple{%d, %s}" % (d, base)
    return base


def test_julia_shim_structs_mirror_the_header():
    cs, js = _c_structs(), _jl_structs()
    names = {"ob_grid_desc": "ObGridDesc", "ob_bc_desc": "ObBcDesc", "ob_closure_desc": "ObClosureDesc", "ob_model_desc": "ObModelDesc"}
    for cname, jname in names.items():
        cf, jf = cs[cname], js[jname]
        assert [f[0] for f in cf] == [f[0] for f in jf], (cname, [f[0] for f in cf], [f[0] for f in jf])
        for (n, ctype, dims), (_, jt) in zip(cf, jf):
            assert _jl_type(ctype, dims, names) == jt, (cname, n, _jl_type(ctype, dims, names), jt)


def _split_args(s):
    """split a Julia argument list at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _call_args(txt, pos):
    """argument string of the call whose '(' is at txt[pos]"""
    depth, i = 0, pos
    while True:
        if txt[i] in "([{":
            depth += 1
        elif txt[i] in ")]}":
            depth -= 1
            if depth == 0:
                return txt[pos + 1:i]
        i += 1


def test_julia_shim_constructor_arities():
    """every positional construction of a descriptor struct passes exactly as many arguments as the struct has fields"""
    import re
    txt = open(JL).read()
    js = _jl_structs()
    for name, fields in js.items():
        calls = [m for m in re.finditer(r"\b%s\(" % name, txt)]
        assert calls, name
        for m in calls:
            args = _split_args(_call_args(txt, m.end() - 1))
            assert len(args) == len(fields), (name, len(args), len(fields), txt[m.start():m.start() + 80])


def test_julia_shim_ccalls_match_the_header():
    """every `@ob name (argtypes) args...` names an exported entry point with that many parameters; every hot-path entry
    point of the header is bound by the shim"""
    import re
    hdr = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"int32_t\s+(ob_\w+)\(([^;]*?)\);", hdr, flags=re.S)}
    txt = open(JL).read()
    used = set()
    for m in re.finditer(r"@ob\(?\s*(ob_\w+),?\s*\(", txt):
        name = m.group(1)
        types = _split_args(_call_args(txt, m.end() - 1))
        assert name in protos, name
        assert len(types) == protos[name], (name, types, protos[name])
        used.add(name)
    for m in re.finditer(r"ccall\(\(:(ob_\w+), lib\)", txt):
        used.add(m.group(1))
    must = {"ob_init", "ob_shutdown", "ob_malloc", "ob_free", "ob_memcpy_h2d", "ob_memcpy_d2h", "ob_memcpy_d2d", "ob_fill", "ob_any_nan", "ob_sync",
            "ob_model_create", "ob_model_destroy", "ob_model_bind_field", "ob_model_set_bc_array", "ob_fill_halo", "ob_fill_halo_array",
            "ob_update_state", "ob_compute_tendencies", "ob_compute_closure_fields", "ob_update_hydrostatic_pressure", "ob_rk3_substep",
            "ob_ab2_step", "ob_cache_tendencies", "ob_compute_pressure_correction", "ob_make_pressure_correction", "ob_time_step_rk3",
            "ob_time_step_ab2", "ob_cell_advection_timescale", "ob_solver_create", "ob_solver_destroy", "ob_poisson_solve",
            "ob_dist_unique_id", "ob_dist_init", "ob_device_count"}
    assert must <= used, sorted(must - used)
