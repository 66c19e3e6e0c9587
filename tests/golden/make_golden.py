"""Regenerate tests/golden/*.npz:  python tests/golden/make_golden.py [name ...]

The reference itself cannot run in this image (no Julia), so these are NOT reference outputs: they are outputs of the CPU
oracle (oracle/, the restatement of the reference CPU() path) frozen at the commit that introduced them.  They pin (i) the
oracle against accidental drift (tests/test_oracle_golden.py, CPU) and (ii) the CUDA path against fixed vectors on the GPU
box, where the oracle is also present but where a simultaneous change of oracle and kernels would otherwise go unseen
(tests/test_gpu_parity.py::test_against_golden_vectors).  Stored per field: every second interior point along each axis,
the sum and the sum of squares of the whole interior.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = {  # name -> (dt, steps)
    "readme_2d": (1e-3, (1, 10)), "ppp_weno5": (1e-3, (1, 10)), "les_amd": (0.5, (1, 10)), "stretched": (0.5, (1, 10)),
    "lilly_bbb": (1e-3, (1,)), "smag_pbp": (1e-3, (1,)), "weno7": (1e-3, (1,)), "centered2_value": (1e-3, (1,)),
    "vi_ppb": (0.05, (1, 10)),
}
SEED = 7


def summarize(om):
    from helpers import oracle_fields, interior_of
    out = {}
    fields = {"u": om.u, "v": om.v, "w": om.w, "pNHS": om.pNHS}
    fields.update(dict(zip(om.tracer_names, om.tracers)))
    for name, arr in oracle_fields(om).items():
        a = np.asarray(interior_of(fields[name], arr), dtype=np.float64)
        out[name + "__sub"] = a[::2, ::2, ::2].copy()
        out[name + "__sum"] = np.array([a.sum(), (a * a).sum()])
    return out


def main():
    from test_gpu_parity import CONFIGS
    only = sys.argv[1:]   # python tests/golden/make_golden.py [name ...]: regenerate a subset
    for name, (dt, snaps) in GOLDEN.items():
        if only and name not in only:
            continue
        cfg = CONFIGS[name]
        om = cfg.oracle_model()
        om.set(**cfg.initial_conditions(SEED))
        store, done = {}, 0
        for n in snaps:
            while done < n:
                om.time_step(dt)
                done += 1
            for k, v in summarize(om).items():
                store["s%d__%s" % (n, k)] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
        print(name, {k: v.shape for k, v in store.items() if k.endswith("__sub") and k.startswith("s1__")})


if __name__ == "__main__":
    main()
