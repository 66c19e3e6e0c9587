"""CPU-only: the conditioning of the ocean-LES configurations with respect to ulp-level input perturbations.

This documents why tests/test_gpu_parity.py holds pNHS of those configurations to a looser bound than u, v, w and the
tracers: perturbing the ORACLE'S OWN initial velocity by one ulp (the size of the difference between two FFT
libraries in the preceding projection) moves pNHS by more than the 1e-11 contract tolerance after one step, while u, v,
w, T, S stay far inside it.  The triply-periodic headline configuration is well conditioned."""
import numpy as np

from helpers import Config, oracle_fields, rel_l2
from test_gpu_parity import CONFIGS


def _run(cfg, perturb, dt):
    om = cfg.oracle_model()
    om.set(**cfg.initial_conditions(3))
    if perturb:
        rng = np.random.default_rng(99)
        for f in (om.u, om.v, om.w):
            f.data[...] = np.nextafter(f.data, f.data + rng.choice([-1.0, 1.0], f.data.shape))
    om.time_step(dt)
    return oracle_fields(om)


def test_les_pressure_is_ill_conditioned_but_velocities_are_not():
    a, b = _run(CONFIGS["les_amd"], False, 0.5), _run(CONFIGS["les_amd"], True, 0.5)
    errs = {k: rel_l2(b[k], a[k]) for k in a}
    assert errs["pNHS"] > 1e-11, errs           # a 1-ulp input change already exceeds the contract tolerance on p
    assert errs["pNHS"] < 1e-9, errs            # ... but stays within the relaxed bound used by the GPU tests
    for k in ("u", "v", "w", "T", "S"):
        assert errs[k] < 1e-12, errs


def test_triply_periodic_config_is_well_conditioned():
    a, b = _run(CONFIGS["ppp_weno5"], False, 1e-3), _run(CONFIGS["ppp_weno5"], True, 1e-3)
    errs = {k: rel_l2(b[k], a[k]) for k in a}
    assert all(e < 1e-12 for e in errs.values()), errs
