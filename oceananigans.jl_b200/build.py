"""Build libocean_b200.so for sm_100a (nvcc cross-compiles without a GPU).

The fused tendency kernels are instantiated per (float type, scheme kind, buffer) in separate
translation units (csrc/tend_inst.cu) so that the build runs in parallel; everything is linked into
ONE shared library that exports the C ABI of include/ocean_b200.h.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libocean_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] + ARCH + os.environ.get("OB_NVCC_EXTRA", "").split()

INSTANCES = [("double", "f64", 0, 0), ("float", "f32", 0, 0)]
INSTANCES += [(t, tn, 1, nb) for t, tn in (("double", "f64"), ("float", "f32")) for nb in (1, 2, 3, 4, 5, 6)]    # Centered 2 .. 12
INSTANCES += [(t, tn, 2, nb) for t, tn in (("double", "f64"), ("float", "f32")) for nb in (2, 3, 4, 5, 6)]       # WENO 3 .. 11


# the staged-ring tendency kernels have their own translation units (csrc/stage_inst.cu): (buffer, mode, closures, LES kind)
# -- the list of csrc/stage_launch.h; a change to tendency_stage.cuh rebuilds only those
STAGE_HEADERS = ("tendency_stage.cuh",)
# headers only ocean_b200.cu includes: a change to them rebuilds neither the tendency nor the staged-kernel units
MAIN_ONLY_HEADERS = ("dist_solver.cuh", "dist.cuh", "kernels.cuh", "streaming.cuh", "poisson.cuh", "diagnostics.cuh", "closures.cuh")
STAGE_VARIANTS = [(3, 0, 0, 0), (3, 0, 1, 0), (3, 1, 1, 2), (3, 1, 1, 3), (3, 1, 2, 2), (3, 1, 2, 3), (3, 2, 1, 2), (3, 2, 1, 3), (3, 2, 2, 2), (3, 2, 2, 3)]


def _headers(stage=True, main=False):
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")) and (stage or f not in STAGE_HEADERS)
          and (main or f not in MAIN_ONLY_HEADERS)]
    hs.append(os.path.join(HERE, "..", "include", "ocean_b200.h"))
    return hs


STAMP = LIB + ".stamp"


def _source_digest():
    import hashlib
    h = hashlib.sha256()
    for path in sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ocean_b200.h"), __file__]):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force=False, verbose=False, jobs=None):
    # the library was linked from exactly these sources (e.g. on the GPU box, where it arrives prebuilt without its object
    # files): the stamp holds the hash of every source taken BEFORE the build that produced the library started
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers(stage=False)
    hdrs_stage = _headers(stage=True)
    jobs_list = []
    objs = []
    main_o = os.path.join(OBJ, "ocean_b200.o")
    src = os.path.join(CSRC, "ocean_b200.cu")
    objs.append(main_o)
    if force or _stale(main_o, [src] + _headers(stage=False, main=True)):
        jobs_list.append([NVCC, *FLAGS, "-c", src, "-o", main_o])
    inst = os.path.join(CSRC, "tend_inst.cu")
    for t, tn, kind, nb in INSTANCES:
        o = os.path.join(OBJ, "tend_%s_k%d_n%d.o" % (tn, kind, nb))
        objs.append(o)
        if force or _stale(o, [inst] + hdrs):
            jobs_list.append([NVCC, *FLAGS, "-DOB_TI_T=%s" % t, "-DOB_TI_TN=%s" % tn, "-DOB_TI_KIND=%d" % kind,
                              "-DOB_TI_NB=%d" % nb, "-c", inst, "-o", o])
    sinst = os.path.join(CSRC, "stage_inst.cu")
    for t, tn in (("double", "f64"), ("float", "f32")):
        for n, mode, ncl, kl in STAGE_VARIANTS:
            o = os.path.join(OBJ, "stage_%s_n%d_m%d_c%d_k%d.o" % (tn, n, mode, ncl, kl))
            objs.append(o)
            if force or _stale(o, [sinst] + hdrs_stage):
                jobs_list.append([NVCC, *FLAGS, "-DOB_SI_T=%s" % t, "-DOB_SI_TN=%s" % tn, "-DOB_SI_N=%d" % n, "-DOB_SI_MODE=%d" % mode,
                                  "-DOB_SI_NCL=%d" % ncl, "-DOB_SI_KL=%d" % kl, "-c", sinst, "-o", o])
    if jobs_list:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for out in ex.map(_run, jobs_list):
                if verbose and out.strip():
                    print(out)
    if jobs_list or not os.path.exists(LIB):
        _run([NVCC, "-shared", *ARCH, "-o", LIB, *objs, "-lcufft", "-lnccl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
    with open(STAMP, "w") as f:
        f.write(digest + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
