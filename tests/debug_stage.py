"""debug aid: where does the staged-ring kernel differ from the marching kernel?  python tests/debug_stage.py [config ...]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocean_b200 as ob
from ocean_b200 import _abi
from test_gpu_parity import CONFIGS
arch = ob.B200(0)
for name in sys.argv[1:] or ["ppp_weno5", "stage_ppp"]:
    cfg = CONFIGS[name]
    d = dict(cfg.__dict__)
    if os.environ.get("DEBUG_FT") == "f32":
        d["ft"] = np.float32
    if os.environ.get("DEBUG_NOCOR"):
        d["coriolis_f"] = None
    if os.environ.get("DEBUG_NOCL"):
        d["closure"] = ()
    from helpers import Config
    cfg = Config(**d)
    bm = cfg.b200_model(arch)
    ob.set(bm, **cfg.initial_conditions(5))
    bm.update_state()
    bm.set_option(_abi.OB_OPT_TENDENCY_KERNEL, 2); bm.compute_tendencies()
    ref = [g.parent() for g in bm.Gn]
    for g in bm.Gn: g.set_parent(np.zeros(g.P[::-1], g.grid.FT))
    bm.set_option(_abi.OB_OPT_TENDENCY_KERNEL, 8); bm.compute_tendencies()
    H = cfg.halo
    for n, (r, g) in enumerate(zip(ref, bm.Gn)):
        got = g.parent()
        bad = np.argwhere(r != got)
        print(name, "field", n, "mismatches", len(bad), "max", float(np.abs(r - got).max()), "rel", float(np.abs(r - got).max() / max(np.abs(r).max(), 1e-300)))
        if len(bad):
            k, j, i = bad[:, 0] - H[2], bad[:, 1] - H[1], bad[:, 2] - H[0]
            print("   lanes", np.bincount(i % 32, minlength=32).tolist())
            print("   rows ", np.bincount(j % 16, minlength=16).tolist())
            print("   k    ", np.bincount(k, minlength=cfg.size[2]).tolist())
