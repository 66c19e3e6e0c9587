"""Build the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/csrc/oracle.c header).

gcc -O2 -ffp-contract=off -fopenmp; the same source is compiled three times (Float64 / Float32 / x87 long double) and
linked into oracle/_build/liboracle.so.  `oracle/_ref/` (a compiled copy of the *reference*) does
not exist for this project: the reference is 100 % Julia (no C/C++ sources under /root/reference),
so there is nothing gcc could build; see DESIGN.md §oracle.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    flags = ["-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-std=gnu11", "-Wall", "-Wno-unused-function", "-Wno-maybe-uninitialized"]
    o64 = os.path.join(OUT_DIR, "oracle_f64.o")
    o32 = os.path.join(OUT_DIR, "oracle_f32.o")
    subprocess.check_call(["gcc", *flags, "-DFT=double", "-c", SRC, "-o", o64])
    subprocess.check_call(["gcc", *flags, "-DFT=float", "-DORACLE_F32", "-c", SRC, "-o", o32])
    o80 = os.path.join(OUT_DIR, "oracle_f80.o")
    subprocess.check_call(["gcc", *flags, "-DFT=long double", "-DORACLE_F80", "-c", SRC, "-o", o80])
    subprocess.check_call(["gcc", "-shared", "-fopenmp", o64, o32, o80, "-lm", "-o", LIB])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
