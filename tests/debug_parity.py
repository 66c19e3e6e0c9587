"""Print per-field parity errors (B200 vs oracle) for a named test configuration: python tests/debug_parity.py les_amd"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import pair, rel_l2  # noqa: E402
from test_gpu_parity import CONFIGS  # noqa: E402
import ocean_b200 as ob  # noqa: E402

arch = ob.B200(0)
for name in sys.argv[1:]:
    om, bm = pair(CONFIGS[name], arch, seed=11)
    g = om.grid
    H, N = g.H, g.N
    sl = (slice(H[2], H[2] + N[2]), slice(H[1], H[1] + N[1]), slice(H[0], H[0] + N[0]))
    print("==", name)
    for n, (og, bg) in enumerate(zip(om.Gn, bm.Gn)):
        a, b = bg.parent()[sl], og.data[sl]
        d = np.abs(a - b)
        k = np.unravel_index(np.argmax(d), d.shape)
        print("G%d rel_l2=%.3e max|d|=%.3e at (k,j,i)=%s |G|max=%.3e" % (n, rel_l2(a, b), d.max(), k, np.abs(b).max()))
    if om.pHY is not None:
        a, b = bm.pressures["pHY"].parent(), om.pHY.data
        print("pHY rel_l2 (parent) = %.3e, max|d| = %.3e" % (rel_l2(a, b), np.abs(a - b).max()))
    for f, nm in ((om.u, "u"), (om.v, "v"), (om.w, "w")):
        a, b = bm.velocities[nm].parent(), f.data
        print(nm, "parent rel_l2 = %.3e" % rel_l2(a, b), "max|d| = %.3e" % np.abs(a - b).max())
    for t, nm in enumerate(om.tracer_names):
        a, b = bm.tracers[nm].parent(), om.tracers[t].data
        print(nm, "parent rel_l2 = %.3e" % rel_l2(a, b), "max|d| = %.3e" % np.abs(a - b).max())
    for m, cf in enumerate(bm.closure_fields):
        if "nue" in cf:
            a, b = cf["nue"].parent(), om.nue[m].data
            print("nue parent rel_l2 = %.3e max|d| = %.3e" % (rel_l2(a, b), np.abs(a - b).max()))
        for t, f in enumerate(cf.get("kappae", [])):
            a, b = f.parent(), om.kappae[m][t].data
            print("kappae", t, "parent rel_l2 = %.3e max|d| = %.3e" % (rel_l2(a, b), np.abs(a - b).max()))
    a, b = bm.pressures["pNHS"].parent(), om.pNHS.data
    print("pNHS parent rel_l2 = %.3e" % rel_l2(a, b))
    from helpers import oracle_fields, b200_fields, interior_of
    dt = 1e-3 if name in ("ppp_weno5", "readme_2d") else 0.5
    for step in range(3):
        om.time_step(dt); ob.time_step(bm, dt)
        of, bf = oracle_fields(om), b200_fields(bm)
        print("step", step + 1, {k: "%.2e" % rel_l2(bf[k], of[k]) for k in of})
