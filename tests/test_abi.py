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
