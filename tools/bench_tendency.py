"""Time the fused tendency kernel alone (CUDA events on the library stream): python tools/bench_tendency.py [N] [reps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocean_b200 as ob  # noqa: E402
from ocean_b200 import _abi  # noqa: E402
from bench import workload_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
arch = ob.B200(0)
fts = {"f64": (np.float64,), "f32": (np.float32,)}.get(os.environ.get("OB_FT", ""), (np.float64, np.float32))
for ft in fts:
    if os.environ.get("OB_CASE") == "les":   # config 3 physics (AMD + ScalarDiffusivity, T, S, FPlane, PPB) on an n^3 grid
        from helpers import Config
        cfg = Config((n, n, n), ((0, float(n)), (0, float(n)), (-n / 2.0, 0.0)), "PPB", advection=("weno", 5), ft=ft,
                     closure=[("amd",), ("scalar", 1.05e-6, 1.46e-7)], buoyancy=("seawater", 9.80665, 2e-4, 8e-4), coriolis_f=1e-4, tracers=("T", "S"),
                     bcs={"u": {"top": ("Flux", -2e-5)}, "T": {"top": ("Flux", 5e-5), "bottom": ("Gradient", 0.005)}, "S": {"top": ("Flux", 5e-8)}})
    else:
        cfg = workload_config(n, ft=ft)
    m = cfg.b200_model(arch)
    ob.set(m, **cfg.initial_conditions(2))
    names = {0: "auto", 1: "generic", 2: "marching", 3: "tma", 8: "stage", 9: "stage-alt"}
    modes = [(int(x), names.get(int(x), "variant%s" % x)) for x in os.environ.get("OB_MODES", "2,8").split(",") if x]
    for mode, name in modes:
        m.set_option(_abi.OB_OPT_TENDENCY_KERNEL, mode)
        for _ in range(3):
            m.compute_tendencies()
        arch.synchronize()
        _abi.call("ob_timer_start", arch.ctx)
        for _ in range(reps):
            m.compute_tendencies()
        ms = C.c_double(0)
        _abi.call("ob_timer_stop", arch.ctx, C.byref(ms))
        t = ms.value / reps
        print("%s %s %d^3: %.3f ms/launch, %.1f GB/s algorithmic (%d B/cell)" % (ft.__name__, name, n, t, 8 * np.dtype(ft).itemsize * n ** 3 / t / 1e6, 8 * np.dtype(ft).itemsize))
    del m
