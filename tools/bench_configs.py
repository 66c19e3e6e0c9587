"""cell-updates/s and per-phase milliseconds for every BASELINE.json configuration that fits one GPU
(python tools/bench_configs.py [scale]); scale < 1 shrinks the grids for a quick check."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Config, stretched_faces  # noqa: E402
import ocean_b200 as ob  # noqa: E402
from ocean_b200 import _abi  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
S = lambda n: max(16, int(round(n * scale / 16)) * 16)
TWO_PI = 2 * np.pi
LES = dict(advection=("weno", 5), closure=[("amd",), ("scalar", 1.05e-6, 1.46e-7)], buoyancy=("seawater", 9.80665, 2e-4, 8e-4),
           coriolis_f=1e-4, tracers=("T", "S"),
           bcs={"u": {"top": ("Flux", -2e-5)}, "T": {"top": ("Flux", 5e-5), "bottom": ("Gradient", 0.005)}, "S": {"top": ("Flux", 5e-8)}})
CASES = [
    ("config1 README 2-D turbulence 128^2 PPF WENO5 F64", Config((128, 128, 1), ((0, TWO_PI), (0, TWO_PI), None), "PPF", advection=("weno", 5)), 0.01, "rk3"),
    ("config2 triply periodic 256^3 WENO5 b ScalarDiffusivity F64", Config((S(256),) * 3, ((0, TWO_PI),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)], buoyancy=("tracer",), tracers=("b",)), 1e-3, "rk3"),
    ("config2 at 512^3 F64", Config((S(512),) * 3, ((0, TWO_PI),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)], buoyancy=("tracer",), tracers=("b",)), 1e-3, "rk3"),
    ("config2 at 256^3 F32", Config((S(256),) * 3, ((0, TWO_PI),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)], buoyancy=("tracer",), tracers=("b",), ft=np.float32), 1e-3, "rk3"),
    ("config3 ocean LES 512^3 PPB AMD+Scalar T,S FPlane F64 RK3", Config((S(512),) * 3, ((0, 512.0), (0, 512.0), (-256.0, 0.0)), "PPB", **LES), 2.0, "rk3"),
    ("config3 ocean LES 512^3 AB2", Config((S(512),) * 3, ((0, 512.0), (0, 512.0), (-256.0, 0.0)), "PPB", timestepper="ab2", **LES), 2.0, "ab2"),
    # classes of models that run the round-1 marching / generic kernels (no staged-ring form): walls in every direction with a
    # Centered scheme, and a higher-order WENO
    ("BBB 256^3 Centered-4 SmagorinskyLilly b F64 (marching kernel, generic bodies)", Config((S(256),) * 3, ((0, 1.0),) * 3, "BBB", advection=("centered", 4), closure=[("lilly", 0.16, 1.0, 1.0)], buoyancy=("tracer",), tracers=("b",)), 1e-3, "rk3"),
    ("PPB 256^3 WENO-7 b ScalarDiffusivity F64 (marching kernel, fast bodies)", Config((S(256),) * 3, ((0, 1.0),) * 3, "PPB", halo=(4, 4, 4), advection=("weno", 7), closure=[("scalar", 1e-3, 1e-3)], buoyancy=("tracer",), tracers=("b",)), 1e-3, "rk3"),
    ("config4 stretched-z 512^2x256 Fourier-tridiagonal F64", Config((S(512), S(512), S(256)), ((0, 512.0), (0, 512.0), stretched_faces(S(256), 256.0)), "PPB", **LES), 2.0, "rk3"),
    ("config4 stretched-z 512^2x256 Fourier-tridiagonal F32", Config((S(512), S(512), S(256)), ((0, 512.0), (0, 512.0), stretched_faces(S(256), 256.0)), "PPB", ft=np.float32, **LES), 2.0, "rk3"),
]
arch = ob.B200(0)
rows = []
for name, cfg, dt, ts in CASES:
    m = cfg.b200_model(arch)
    rng = np.random.default_rng(3)
    g = m.grid
    ic = {}
    for f, a in (("u", 1e-3), ("v", 0.0), ("w", 1e-3)):
        if cfg.topology["uvw".index(f)] == "F":
            continue
        fld = m.velocities[f]
        ic[f] = (a * rng.standard_normal(fld.n[::-1]) if "T" in cfg.tracers else 0.1 * rng.uniform(-1, 1, fld.n[::-1])).astype(cfg.ft)
    zc = g.nodes(2, "c")[:, None, None] if cfg.topology[2] != "F" else 0.0
    for t in cfg.tracers:
        base = {"b": 1.0 * zc, "T": 20 + 0.005 * zc, "S": 35.0 + 0 * zc}[t]
        ic[t] = (base + 1e-3 * rng.uniform(-1, 1, m.tracers[t].n[::-1])).astype(cfg.ft)
    ob.set(m, **ic)
    for _ in range(3):
        ob.time_step(m, dt)
    _abi.call("ob_reset_timing", m.handle); _abi.call("ob_enable_timing", m.handle, 1)
    steps = 5
    arch.synchronize()
    _abi.call("ob_timer_start", arch.ctx)
    for _ in range(steps):
        ob.time_step(m, dt)
    ms = C.c_double(0)
    _abi.call("ob_timer_stop", arch.ctx, C.byref(ms))
    nph = C.c_int32(0); _abi.call("ob_phase_count", C.byref(nph))
    ph = {}
    for p in range(nph.value):
        tt, cc = C.c_double(0), C.c_int64(0)
        _abi.call("ob_phase_time_ms", m.handle, p, C.byref(tt), C.byref(cc))
        ph[_abi.lib().ob_phase_name(p).decode()] = round(tt.value / steps, 3)
    cells = g.N[0] * g.N[1] * g.N[2]
    ok = not m.velocities["u"].any_nan()
    row = {"config": name, "grid": list(g.N), "ms_per_step": round(ms.value / steps, 3), "Gcell_updates_per_s": round(cells * steps / ms.value / 1e6, 3),
           "phases_ms_per_step": ph, "finite": ok}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del m
