"""Full-size (BASELINE.json configs[1], 256^3 Float64) property tests through the C ABI: size-independent invariants
instead of an oracle comparison (the CPU oracle needs minutes at this size)."""
import ctypes as C

import numpy as np
import pytest

from helpers import Config, rel_l2

pytestmark = pytest.mark.gpu
TWO_PI = 2 * np.pi


@pytest.fixture(scope="module")
def big(arch):
    import ocean_b200 as ob
    cfg = Config((256, 256, 256), ((0, TWO_PI),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)],
                 buoyancy=("tracer",), tracers=("b",))
    m = cfg.b200_model(arch)
    ob.set(m, **cfg.initial_conditions(2))
    return cfg, m


def _divergence(m):
    u, v, w = (m.velocities[n].interior().astype(np.float64) for n in "uvw")
    g = m.grid
    dx, dy, dz = float(g.dF[0]), float(g.dF[1]), float(g.dF[2])
    return (np.roll(u, -1, 2) - u) / dx + (np.roll(v, -1, 1) - v) / dy + (np.roll(w, -1, 0) - w) / dz


def test_projection_and_conservation_at_256(big):
    """after set! and after two RK3 steps: max|div u| < 5e-8 (test_time_stepping.jl:545-573) and the tracer integral
    is conserved (periodic domain, :179-213)"""
    import ocean_b200 as ob
    cfg, m = big
    assert np.abs(_divergence(m)).max() < 5e-8
    b0 = m.tracers["b"].interior().astype(np.float64).sum()
    for _ in range(2):
        ob.time_step(m, 1e-3)
    assert np.abs(_divergence(m)).max() < 5e-8
    b1 = m.tracers["b"].interior().astype(np.float64).sum()
    assert abs(b1 - b0) <= 1e-12 * abs(b0) + 1e-6
    assert not m.velocities["u"].any_nan()


def test_halo_fill_is_idempotent_and_periodic_at_256(big):
    cfg, m = big
    f = m.tracers["b"]
    m.fill_halo_regions("b")
    a = f.parent()
    m.fill_halo_regions("b")
    assert np.array_equal(a, f.parent())
    H, N = m.grid.H[0], m.grid.N[0]
    assert np.array_equal(a[:, :, :H], a[:, :, N:N + H]) and np.array_equal(a[:, :, N + H:], a[:, :, H:2 * H])
    assert np.array_equal(a[:H], a[N:N + H]) and np.array_equal(a[:, N + H:], a[:, H:2 * H])


def test_marching_and_generic_tendency_kernels_agree_at_256(big):
    from ocean_b200 import _abi
    cfg, m = big
    m.set_option(_abi.OB_OPT_TENDENCY_KERNEL, 1)
    m.compute_tendencies()
    ref = [g.interior() for g in m.Gn]
    m.set_option(_abi.OB_OPT_TENDENCY_KERNEL, 2)
    m.compute_tendencies()
    for r, g in zip(ref, m.Gn):
        assert rel_l2(g.interior(), r) <= 1e-13
    march = [g.interior() for g in m.Gn]
    m.set_option(_abi.OB_OPT_TENDENCY_KERNEL, 3)   # TMA-staged planes: same arithmetic, bit-identical
    m.compute_tendencies()
    for r, g in zip(march, m.Gn):
        assert np.array_equal(g.interior(), r)
    for mode in (8, 0):   # staged-ring kernel (explicitly and as the automatic choice): same arithmetic, bit-identical
        for g in m.Gn:
            g.set_parent(np.zeros(g.P[::-1], g.grid.FT))
        m.set_option(_abi.OB_OPT_TENDENCY_KERNEL, mode)
        m.compute_tendencies()
        for r, g in zip(march, m.Gn):
            assert np.array_equal(g.interior(), r)


def test_poisson_solver_inverts_the_laplacian_at_256(arch):
    import ocean_b200 as ob
    cfg = Config((256, 256, 256), ((0, TWO_PI),) * 3, "PPP")
    solver = ob.FFTBasedPoissonSolver(cfg.b200_grid(arch))
    rng = np.random.default_rng(9)
    phi = rng.standard_normal((256, 256, 256))
    phi -= phi.mean()
    d = TWO_PI / 256
    lap = sum((np.roll(phi, -1, ax) - 2 * phi + np.roll(phi, 1, ax)) for ax in range(3)) / d ** 2
    got = ob.solve(solver, lap)
    assert rel_l2(got, phi) <= 1e-11
