"""Time compute_closure_fields (AMD) alone at n^3 for the config-3 physics: python tools/bench_closure.py [n] [reps]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ocean_b200 as ob
from ocean_b200 import _abi
from helpers import Config
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
arch = ob.B200(0)
for ft in (np.float64, np.float32):
    cfg = Config((n, n, n), ((0, float(n)), (0, float(n)), (-n / 2.0, 0.0)), "PPB", advection=("weno", 5), ft=ft,
                 closure=[("amd",), ("scalar", 1.05e-6, 1.46e-7)], buoyancy=("seawater", 9.80665, 2e-4, 8e-4), coriolis_f=1e-4, tracers=("T", "S"))
    m = cfg.b200_model(arch)
    ob.set(m, **cfg.initial_conditions(2))
    for _ in range(3):
        _abi.call("ob_compute_closure_fields", m.handle)
    arch.synchronize()
    _abi.call("ob_timer_start", arch.ctx)
    for _ in range(reps):
        _abi.call("ob_compute_closure_fields", m.handle)
    ms = C.c_double(0)
    _abi.call("ob_timer_stop", arch.ctx, C.byref(ms))
    t = ms.value / reps
    w = np.dtype(ft).itemsize
    print("%s amd closure fields %d^3: %.3f ms/launch, %.0f GB/s algorithmic (8 words/cell)" % (ft.__name__, n, t, 8 * w * n ** 3 / t / 1e6))
    del m
