"""CPU: host-side mirror logic that needs no device -- boundary-condition regularisation (field_boundary_conditions.jl:15-36),
descriptor contents, input validation (input_validation.jl:71-93), schedules, clock bookkeeping."""
import numpy as np
import pytest


class _NoDevice:
    ctx = None


def _grid(topology="PPB", size=(8, 6, 4), halo=None, ft=np.float64, z=None):
    import ocean_b200 as ob
    T = {"P": ob.Periodic, "B": ob.Bounded, "F": ob.Flat}
    topo = tuple(T[t] for t in topology)
    nf = [n for n, t in zip(size, topology) if t != "F"]
    kw = dict(size=tuple(nf), topology=topo)
    if halo is not None:
        kw["halo"] = halo
    ext = {"x": (0, 1.0), "y": (0, 2.0), "z": (-1.0, 0.0) if z is None else z}
    for name, t in zip("xyz", topology):
        if t != "F":
            kw[name] = ext[name]
    return ob.RectilinearGrid(_NoDevice(), ft, **kw)


def test_default_boundary_conditions_follow_the_reference():
    """Periodic -> PBC; Bounded + Center -> no-flux; Bounded + Face (normal velocity) -> impenetrable; Flat -> nothing"""
    from ocean_b200.fields import regularize_bcs
    g = _grid("PBB")
    u, v, w, c = (regularize_bcs(g, loc) for loc in ("fcc", "cfc", "ccf", "ccc"))
    assert u["west"].kind == u["east"].kind == "Periodic"
    assert u["south"].kind == "Flux" and u["south"].value is None and u["top"].kind == "Flux"
    assert v["south"].kind == v["north"].kind == "Impenetrable" and v["bottom"].kind == "Flux"
    assert w["bottom"].kind == w["top"].kind == "Impenetrable"
    assert all(c[s].kind == "Flux" for s in ("south", "north", "bottom", "top"))
    g2 = _grid("PPF", size=(8, 8, 1))
    assert regularize_bcs(g2, "ccc")["top"] is None and regularize_bcs(g2, "ccf")["bottom"] is None


def test_user_boundary_conditions_and_validation():
    import ocean_b200 as ob
    from ocean_b200.fields import regularize_bcs, bc_desc
    from ocean_b200 import _abi
    g = _grid("PPB")
    bcs = regularize_bcs(g, "ccc", ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(2.5), bottom=ob.GradientBoundaryCondition(-0.1)))
    d = bc_desc(bcs)
    assert list(d.kind) == [_abi.OB_BC_PERIODIC] * 4 + [_abi.OB_BC_GRADIENT, _abi.OB_BC_FLUX] and d.value[5] == 2.5 and d.value[4] == -0.1
    arr = np.ones((6, 8))
    d2 = bc_desc(regularize_bcs(g, "ccc", ob.FieldBoundaryConditions(top=ob.ValueBoundaryCondition(arr))))
    assert d2.kind[5] == _abi.OB_BC_VALUE and d2.value[5] == 0.0           # arrays travel through ob_model_set_bc_array
    with pytest.raises(ValueError):
        regularize_bcs(g, "ccc", ob.FieldBoundaryConditions(west=ob.FluxBoundaryCondition(1.0)))   # Periodic direction
    with pytest.raises(ob.OceanB200Error):
        regularize_bcs(g, "ccc", ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(lambda x, y, t: 1.0)))
    with pytest.raises(ValueError):
        ob.FieldBoundaryConditions(upper=ob.FluxBoundaryCondition(1.0))
    with pytest.raises(ob.OceanB200Error):
        ob.OpenBoundaryCondition(1.0)


def test_grid_input_validation_and_descriptor():
    import ocean_b200 as ob
    from ocean_b200 import _abi
    g = _grid("PPB", size=(8, 6, 4))
    assert g.H == (3, 3, 3) and (g.Nx, g.Ny, g.Nz) == (8, 6, 4)          # default halo (3,3,3)
    assert _grid("PPB", size=(2, 6, 4)).H == (2, 3, 3)                    # clipped to the size (validate_halo)
    d = g.desc()
    assert list(d.N) == [8, 6, 4] and list(d.topology) == [_abi.OB_PERIODIC, _abi.OB_PERIODIC, _abi.OB_BOUNDED]
    assert d.d[0] == 1.0 / 8 and d.d[1] == 2.0 / 6 and d.d[2] == 0.25 and not d.dzf_host
    gs = _grid("PPB", size=(8, 6, 4), z=np.array([-1.0, -0.6, -0.3, -0.1, 0.0]))
    ds = gs.desc()
    assert ds.d[2] == 0.0 and ds.n_dzf == 4 + 2 * 3 + 1 and ds.n_dzc == 4 + 2 * 3   # Δᶠ: Nz+2Hz+1 entries, Δᶜ: Nz+2Hz (ocean_b200.h)
    with pytest.raises(ValueError):
        _grid("PPB", size=(8, 6, 4), halo=(9, 1, 1))                      # halo > size
    with pytest.raises(ValueError):
        _grid("PPB", size=(8, 6, 4), z=(0.0, -1.0))                       # decreasing interval
    with pytest.raises(ValueError):
        _grid("PPB", size=(8, 6, 4), z=np.array([-1.0, -0.6, -0.7, -0.1, 0.0]))   # non-monotone faces
    with pytest.raises(ValueError):
        ob.RectilinearGrid(None, size=(4, 4, 4), extent=(1, 1, 1))        # an architecture is required
    with pytest.raises(NotImplementedError):
        ob.CPU()                                                          # no CPU path in this library
    with pytest.raises(ob.OceanB200Error):
        T = (ob.Periodic, ob.Periodic, ob.Bounded)
        ob.RectilinearGrid(_NoDevice(), size=(4, 4, 4), x=np.linspace(0, 1, 5) ** 2, y=(0, 1), z=(0, 1), topology=T).desc()   # only z may stretch
    flat = _grid("PPF", size=(8, 8, 1))
    assert flat.N == (8, 8, 1) and flat.H == (3, 3, 0) and flat.desc().topology[2] == _abi.OB_FLAT
    wh = g.with_halo((4, 4, 4))
    assert wh.H == (4, 4, 4) and wh.N == g.N and np.array_equal(wh.nodes(2, "c"), g.nodes(2, "c"))


def test_schedules_and_wizard_arithmetic():
    import ocean_b200 as ob

    class _M:
        class clock:
            iteration, time = 0, 0.0
    m = _M()
    s = ob.IterationInterval(3)
    fired = []
    for it in range(7):
        m.clock.iteration = it
        if s(m):
            fired.append(it)
    assert fired == [0, 3, 6]
    w = ob.TimeStepWizard(cfl=0.5, max_change=1.1, min_change=0.5, max_dt=10.0)

    class _Model:
        def cell_advection_timescale(self):
            return 4.0
    assert w.new_time_step(1.0, _Model()) == pytest.approx(1.1)           # limited by max_change (wizard.jl)
    assert w.new_time_step(10.0, _Model()) == pytest.approx(5.0)          # limited by min_change
    assert ob.TimeStepWizard(cfl=0.5, max_dt=1.5).new_time_step(1.4, _Model()) == pytest.approx(1.5)


def test_time_interval_schedule_follows_the_reference():
    """Utils/schedules.jl: the first actuation time is the clock time at initialisation (fires there), later actuations are at
    first + n * interval without accumulated rounding; aligned_time_step clips dt to the next actuation (run.jl:43-55)"""
    import ocean_b200 as ob
    from ocean_b200 import simulations as S

    class _Clock:
        iteration, time = 0, 0.0

    class _M:
        clock = _Clock()
    m = _M()
    s = ob.TimeInterval(0.3)
    fired = []
    for it in range(12):
        m.clock.time = 0.1 * it
        if s(m):
            fired.append(it)
    assert fired == [0, 3, 6, 9]
    # restored at t > 0: fires at initialisation, then only every interval (not on every iteration until caught up)
    m.clock.time = 5.0
    s2 = ob.TimeInterval(1.0)
    assert s2(m)
    m.clock.time = 5.1
    assert not s2(m)
    m.clock.time = 6.0
    assert s2(m)
    assert s2.next_actuation_time() == pytest.approx(7.0)

    class _Sim:
        stop_time = float("inf")
        model = m
        callbacks = {}
        output_writers = {"w": type("W", (), {"schedule": s2})()}
    m.clock.time = 6.75
    assert S._aligned_time_step(_Sim(), 1.0) == pytest.approx(0.25)


def test_closure_constructors_take_a_time_discretization_like_the_reference():
    """Smagorinsky([time_discretization]; coefficient, Pr), AnisotropicMinimumDissipation([time_discretization]; C, Cν, Cκ, Cb),
    ScalarDiffusivity([time_discretization]; ν, κ) (smagorinsky.jl:76-84, anisotropic_minimum_dissipation.jl:124-139,
    scalar_diffusivity.jl:113-137): an instance or the type itself, first positional argument; anything else is a TypeError"""
    import ocean_b200 as ob
    VI, EX = ob.VerticallyImplicitTimeDiscretization, ob.ExplicitTimeDiscretization
    for make in (lambda *a: ob.Smagorinsky(*a, coefficient=0.2, Pr=2.0), lambda *a: ob.SmagorinskyLilly(*a, C=0.2, Cb=1.0),
                 lambda *a: ob.AnisotropicMinimumDissipation(*a, C=0.25), lambda *a: ob.ScalarDiffusivity(*a, nu=1e-3, kappa=1e-4)):
        assert make().vertically_implicit is False
        assert make(EX()).vertically_implicit is False
        assert make(VI()).vertically_implicit is True
        assert make(VI).vertically_implicit is True
        with pytest.raises(TypeError):
            make("implicit")
    s = ob.SmagorinskyLilly(VI(), C=0.2, Cb=0.5, Pr=3.0)
    assert (s.cs, s.cb, s.lilly, s.Pr) == (0.2, 0.5, True, 3.0)
    a = ob.AnisotropicMinimumDissipation(Cν=0.3, Cκ=0.4)
    assert (a.Cnu, a.Ckappa, a.Cb) == (0.3, 0.4, None)


def test_advection_scheme_orders():
    """every buffer the reference builds (src/Advection/Advection.jl:52): WENO 3 .. 11 (odd), Centered 2 .. 12 (even)"""
    import ocean_b200 as ob
    assert [ob.WENO(order=o).buffer for o in (3, 5, 7, 9, 11)] == [2, 3, 4, 5, 6]
    assert [ob.Centered(order=o).buffer for o in (2, 4, 6, 8, 10, 12)] == [1, 2, 3, 4, 5, 6]
    for bad in (4, 1):
        with pytest.raises(ValueError):
            ob.WENO(order=bad)
    with pytest.raises(ValueError):
        ob.Centered(order=3)


def test_dynamic_smagorinsky_constructor():
    """DynamicSmagorinsky(; averaging, Pr, schedule, minimum_numerator) (dynamic_coefficient.jl:107-118): the averaging
    dimensions as an integer, a tuple or a colon; LagrangianAveraging (the reference default) is representable but rejected
    when a model is built from it"""
    import ocean_b200 as ob
    assert ob.DynamicSmagorinsky(averaging=1).dynamic["averaging"] == (1,)
    assert ob.DynamicSmagorinsky(averaging=(1, 2), Pr=2.0).dynamic["averaging"] == (1, 2)
    assert ob.DynamicSmagorinsky(averaging="colon").dynamic["averaging"] == (1, 2, 3)
    assert ob.DynamicSmagorinsky(averaging=slice(None)).dynamic["averaging"] == (1, 2, 3)
    d = ob.DynamicSmagorinsky()
    assert isinstance(d.dynamic["averaging"], ob.LagrangianAveraging) and d.dynamic["minimum_numerator"] == 1e-32
    assert ob.DynamicSmagorinsky(ob.VerticallyImplicitTimeDiscretization(), averaging=(1, 2)).vertically_implicit is True
    assert ob.Smagorinsky().dynamic is None and ob.SmagorinskyLilly().dynamic is None
    with pytest.raises(ValueError):
        ob.DynamicSmagorinsky(averaging=(1, 4))
    with pytest.raises(ValueError):
        ob.DynamicSmagorinsky(averaging=())
