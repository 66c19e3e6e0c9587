"""A few RK3 steps of the ocean-LES physics (config 3: AMD + ScalarDiffusivity, T, S, FPlane, PPB; `stretched`: config 4's
stretched z and Fourier-tridiagonal solver) for profiler captures: python tools/les_step.py [n] [regular|stretched]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Config, stretched_faces  # noqa: E402
import ocean_b200 as ob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
stretched = len(sys.argv) > 2 and sys.argv[2] == "stretched"
nz = n // 2 if stretched else n
z = stretched_faces(nz, float(nz)) if stretched else (-float(nz), 0.0)
cfg = Config((n, n, nz), ((0, float(n)), (0, float(n)), z), "PPB", advection=("weno", 5),
             closure=[("amd",), ("scalar", 1.05e-6, 1.46e-7)], buoyancy=("seawater", 9.80665, 2e-4, 8e-4), coriolis_f=1e-4, tracers=("T", "S"),
             bcs={"u": {"top": ("Flux", -2e-5)}, "T": {"top": ("Flux", 5e-5), "bottom": ("Gradient", 0.005)}, "S": {"top": ("Flux", 5e-8)}})
arch = ob.B200(0)
m = cfg.b200_model(arch)
ob.set(m, **cfg.initial_conditions(2))
for _ in range(4):
    ob.time_step(m, 0.5)
arch.synchronize()
print("done", float(np.abs(m.velocities["u"].interior()).max()))
