"""Host-side mirror of `NonhydrostaticModel(grid; advection, closure, buoyancy, coriolis, tracers, timestepper,
boundary_conditions)`, `time_step!`, `set!`, `update_state!` -- every device operation is one call into the C ABI.

Reference: src/Models/NonhydrostaticModels/nonhydrostatic_model.jl:124-313 (constructor: halo inflation :318-332,
pHY′ only when buoyancy is set and z is not Periodic :177-198), set_nonhydrostatic_model.jl:39-74,
src/TimeSteppers/runge_kutta_3.jl:103-184, quasi_adams_bashforth_2.jl:90-126, clock.jl:147-173.
Anything the C ABI cannot express (function-valued BCs/forcings, background fields, immersed boundaries, other
closures / equations of state) raises OceanB200Error(OB_ERR_UNSUPPORTED) -- there is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .fields import Field, regularize_bcs, bc_desc, FieldBoundaryConditions
from .grids import RectilinearGrid


# ---- schemes / closures / buoyancy / coriolis (plain descriptors, as in Julia) ---------------------------------
class Centered:
    def __init__(self, order=2):
        if order % 2 or order < 2:
            raise ValueError("Centered reconstruction scheme is defined only for even orders")
        self.order, self.buffer = order, order // 2


class WENO:
    """WENO(; order=5).  `weight_computation` in {None, 'NormalDivision', 'BackendOptimizedDivision'}: None takes
    `default_weno_weight_computation(::B200)` = BackendOptimizedDivision, i.e. rcp.approx + one cubic Newton step --
    the construction the reference's own CUDA backend uses (ext/OceananigansCUDAExt.jl:147-163; materialize_advection.jl:
    45-51); 'NormalDivision' selects IEEE division (OB_DIV_EXACT).  Both stay within the parity tolerances."""

    def __init__(self, order=5, weight_computation=None):
        if order % 2 == 0 or order < 3:
            raise ValueError("WENO reconstruction scheme is defined only for odd orders >= 3")
        self.order, self.buffer = order, (order + 1) // 2
        self.weight_computation = weight_computation


class ExplicitTimeDiscretization:
    pass


class VerticallyImplicitTimeDiscretization:
    """ScalarDiffusivity(VerticallyImplicitTimeDiscretization(), ν=..., κ=...): vertical diffusion is taken out of the
    explicit tendencies and solved implicitly after every substep (vertically_implicit_diffusion_solver.jl)"""


class ScalarDiffusivity:
    """ScalarDiffusivity([time_discretization,] ν, κ) (scalar_diffusivity.jl:113-137)"""

    def __init__(self, time_discretization=None, nu=0.0, kappa=0.0, ν=None, κ=None):
        if isinstance(time_discretization, type):
            time_discretization = time_discretization()
        if time_discretization is not None and not isinstance(time_discretization, (ExplicitTimeDiscretization, VerticallyImplicitTimeDiscretization)):
            raise TypeError("time_discretization must be ExplicitTimeDiscretization() or VerticallyImplicitTimeDiscretization()")
        self.time_discretization = time_discretization or ExplicitTimeDiscretization()
        self.vertically_implicit = isinstance(self.time_discretization, VerticallyImplicitTimeDiscretization)
        self.nu = nu if ν is None else ν
        self.kappa = kappa if κ is None else κ
        self.required_halo = 1


def _time_discretization(td):
    if isinstance(td, type):
        td = td()
    if td is not None and not isinstance(td, (ExplicitTimeDiscretization, VerticallyImplicitTimeDiscretization)):
        raise TypeError("time_discretization must be ExplicitTimeDiscretization() or VerticallyImplicitTimeDiscretization()")
    return td or ExplicitTimeDiscretization()


class Smagorinsky:
    """Smagorinsky([time_discretization]; coefficient, Pr) (smagorinsky.jl:76-84)"""

    def __init__(self, time_discretization=None, coefficient=0.16, Pr=1.0):
        self.time_discretization = _time_discretization(time_discretization)
        self.vertically_implicit = isinstance(self.time_discretization, VerticallyImplicitTimeDiscretization)
        self.cs, self.Pr, self.lilly, self.cb = coefficient, Pr, False, 0.0
        self.dynamic = None   # DynamicSmagorinsky(): {"averaging": dims, "schedule": ..., "minimum_numerator": ...}
        self.required_halo = 2


class LagrangianAveraging:
    """averaging = LagrangianAveraging() of DynamicCoefficient: outside the B200 hot path (rejected at model construction)"""


def DynamicSmagorinsky(time_discretization=None, averaging=None, Pr=1.0, schedule=None, minimum_numerator=1e-32):
    """DynamicSmagorinsky([time_discretization]; averaging, Pr, schedule, minimum_numerator) (dynamic_coefficient.jl:107-118): a
    Smagorinsky closure whose coefficient is computed from the flow (Bou-Zeid et al. 2005, scale-invariant).  `averaging`: an
    integer or a tuple of integers (dimensions averaged over), or `slice(None)` / "colon" for all three; the reference default,
    LagrangianAveraging(), is not on the B200 hot path.  Only the default schedule (every iteration) is supported."""
    s = Smagorinsky(time_discretization, coefficient=0.0, Pr=Pr)
    if averaging is None or isinstance(averaging, LagrangianAveraging):
        averaging = LagrangianAveraging()
    elif isinstance(averaging, int):
        averaging = (averaging,)
    elif averaging == "colon" or averaging == slice(None):
        averaging = (1, 2, 3)
    else:
        averaging = tuple(int(d) for d in averaging)
    if not isinstance(averaging, LagrangianAveraging) and (not averaging or any(d not in (1, 2, 3) for d in averaging)):
        raise ValueError("averaging dimensions must be a subset of (1, 2, 3)")
    s.dynamic = {"averaging": averaging, "schedule": schedule, "minimum_numerator": minimum_numerator}
    return s


def SmagorinskyLilly(time_discretization=None, C=0.16, Cb=1.0, Pr=1.0):
    s = Smagorinsky(time_discretization, coefficient=C, Pr=Pr)
    s.lilly, s.cb = True, Cb
    return s


class AnisotropicMinimumDissipation:
    """AnisotropicMinimumDissipation([time_discretization]; C, Cν, Cκ, Cb) (anisotropic_minimum_dissipation.jl:124-139)"""

    def __init__(self, time_discretization=None, C=1.0 / 3.0, Cnu=None, Ckappa=None, Cb=None, Cν=None, Cκ=None):
        self.time_discretization = _time_discretization(time_discretization)
        self.vertically_implicit = isinstance(self.time_discretization, VerticallyImplicitTimeDiscretization)
        Cnu = Cν if Cν is not None else Cnu
        Ckappa = Cκ if Cκ is not None else Ckappa
        self.Cnu = C if Cnu is None else Cnu
        self.Ckappa = C if Ckappa is None else Ckappa
        self.Cb = Cb
        self.required_halo = 2


class BuoyancyTracer:
    pass


class LinearEquationOfState:
    def __init__(self, thermal_expansion=1.67e-4, haline_contraction=7.80e-4):
        self.thermal_expansion, self.haline_contraction = thermal_expansion, haline_contraction


class SeawaterBuoyancy:
    def __init__(self, gravitational_acceleration=9.80665, equation_of_state=None):
        self.g = gravitational_acceleration
        self.equation_of_state = equation_of_state or LinearEquationOfState()
        if not isinstance(self.equation_of_state, LinearEquationOfState):
            raise _abi.OceanB200Error(-3, "only LinearEquationOfState is on the B200 hot path")


class FPlane:
    def __init__(self, f=None, rotation_rate=7.292115e-5, latitude=None):
        if f is None:
            if latitude is None:
                raise ValueError("FPlane needs f or latitude")
            f = 2 * rotation_rate * math.sin(math.radians(latitude))
        self.f = f


class Clock:
    """clock.jl:19-25"""

    def __init__(self):
        self.time, self.iteration, self.stage = 0.0, 0, 1
        self.last_dt, self.last_stage_dt = math.inf, math.inf


def _per_tracer(v, names, t):
    if isinstance(v, dict):
        return v[names[t]]
    if isinstance(v, (tuple, list)):
        return v[t]
    return v


class NonhydrostaticModel:
    def __init__(self, grid, *, advection=None, closure=None, buoyancy=None, coriolis=None, tracers=(),
                 timestepper="RungeKutta3", boundary_conditions=None, forcing=None, background_fields=None,
                 stokes_drift=None, particles=None, biogeochemistry=None, free_surface=None, chi=0.1):
        for name, val in (("forcing", forcing), ("background_fields", background_fields), ("stokes_drift", stokes_drift),
                          ("particles", particles), ("biogeochemistry", biogeochemistry), ("free_surface", free_surface)):
            if val:
                raise _abi.OceanB200Error(-3, "%s is outside the B200 hot path (SURVEY.md §2): no CPU fallback" % name)
        if not isinstance(grid, RectilinearGrid):
            raise _abi.OceanB200Error(-3, "only RectilinearGrid is supported")
        advection = Centered() if advection is None else advection
        self.advection = advection
        self.closures = [] if closure is None else (list(closure) if isinstance(closure, (tuple, list)) else [closure])
        self.buoyancy, self.coriolis = buoyancy, coriolis
        self.tracer_names = (tracers,) if isinstance(tracers, str) else tuple(tracers)
        nt = len(self.tracer_names)
        if nt > _abi.OB_MAX_TRACERS or len(self.closures) > _abi.OB_MAX_CLOSURES:
            raise _abi.OceanB200Error(-3, "too many tracers / closures")
        ts = timestepper.lstrip(":")
        if ts not in ("RungeKutta3", "QuasiAdamsBashforth2"):
            raise ValueError("timestepper = :%s is not supported (RungeKutta3, QuasiAdamsBashforth2)" % ts)
        self.timestepper = ts
        # inflate_grid_halo_size (nonhydrostatic_model.jl:318-332)
        need = max([getattr(advection, "buffer", 1)] + [c.required_halo for c in self.closures] + [1])
        req = tuple(0 if t == _abi.OB_FLAT else max(h, need) for h, t in zip(grid.H, grid.topo))
        if any(r > h for r, h in zip(req, grid.H)):
            grid = grid.with_halo(req)
        self.grid = g = grid
        self.architecture = arch = g.architecture
        FT = g.FT
        self.chi = FT(chi)
        user = dict(boundary_conditions or {})
        for k in user:
            if k not in ("u", "v", "w") + self.tracer_names:
                raise ValueError("boundary_conditions key %r is not a prognostic field" % k)

        def mk(name, loc, aux=False):
            return Field(g, loc, regularize_bcs(g, loc, user.get(name), auxiliary=aux), name)

        self.velocities = {"u": mk("u", "fcc"), "v": mk("v", "cfc"), "w": mk("w", "ccf")}
        self.tracers = {n: mk(n, "ccc") for n in self.tracer_names}
        self.pressures = {"pNHS": mk("pNHS", "ccc")}
        has_phy = buoyancy is not None and g.topo[2] != _abi.OB_PERIODIC
        if has_phy:
            self.pressures["pHY"] = mk("pHY", "ccc")
        names = ("u", "v", "w") + self.tracer_names
        locs = ("fcc", "cfc", "ccf") + ("ccc",) * nt
        self.Gn = [Field(g, l, regularize_bcs(g, l, None, auxiliary=True), "Gn_" + n) for n, l in zip(names, locs)]
        self.Gm = [Field(g, l, regularize_bcs(g, l, None, auxiliary=True), "Gm_" + n) for n, l in zip(names, locs)]
        self.closure_fields = []
        for m, c in enumerate(self.closures):
            cf = {}
            if not isinstance(c, ScalarDiffusivity):
                cf["nue"] = mk("nue%d" % m, "ccc")
            if isinstance(c, AnisotropicMinimumDissipation):
                cf["kappae"] = [mk("kappae%d_%s" % (m, n), "ccc") for n in self.tracer_names]
            self.closure_fields.append(cf)
        self.clock = Clock()

        # ---- descriptor ------------------------------------------------------------------------------------------
        d = _abi.ModelDesc()
        d.grid = g.desc()
        self._keepalive = (g.dF, g.dC)
        if isinstance(advection, WENO):
            d.advection_kind, d.advection_order = _abi.OB_ADV_WENO, advection.order
            wc = advection.weight_computation
            d.weno_division = _abi.OB_DIV_EXACT if wc == "NormalDivision" else _abi.OB_DIV_RCP_NEWTON
        elif isinstance(advection, Centered):
            d.advection_kind, d.advection_order = _abi.OB_ADV_CENTERED, advection.order
        else:
            raise _abi.OceanB200Error(-3, "advection scheme %r is outside the B200 hot path" % (advection,))
        d.n_closures = len(self.closures)
        for m, c in enumerate(self.closures):
            cd = d.closures[m]
            if isinstance(c, ScalarDiffusivity):
                cd.kind, cd.nu = _abi.OB_CLOSURE_SCALAR_DIFFUSIVITY, float(FT(c.nu))
                cd.vertically_implicit = int(c.vertically_implicit)
                if c.vertically_implicit and g.topo[2] != _abi.OB_BOUNDED:
                    raise ValueError("VerticallyImplicitTimeDiscretization can only be specified on grids that are Bounded in the z-direction.")
                for t in range(nt):
                    kv = _per_tracer(c.kappa, self.tracer_names, t)
                    if callable(kv) or callable(c.nu):
                        raise _abi.OceanB200Error(-3, "function-valued diffusivities cannot cross the C ABI")
                    cd.kappa[t] = float(FT(kv))
            elif isinstance(c, Smagorinsky):
                cd.kind, cd.cs, cd.lilly, cd.cb = _abi.OB_CLOSURE_SMAGORINSKY, float(FT(c.cs)), int(c.lilly), float(FT(c.cb))
                cd.vertically_implicit = int(c.vertically_implicit)
                if c.dynamic is not None:
                    if isinstance(c.dynamic["averaging"], LagrangianAveraging):
                        raise _abi.OceanB200Error(-3, "DynamicSmagorinsky with LagrangianAveraging is outside the B200 hot path "
                                                      "(directional averaging, e.g. averaging=(1, 2), is supported)")
                    if c.dynamic["schedule"] is not None:
                        raise _abi.OceanB200Error(-3, "DynamicCoefficient schedules other than IterationInterval(1) are not supported")
                    cd.dynamic = 1
                    cd.averaging_dims = sum(1 << (d - 1) for d in {*c.dynamic["averaging"]})   # (`set` is set!(model; ...) in this module)
                    cd.minimum_numerator = float(FT(c.dynamic["minimum_numerator"]))
                for t in range(nt):
                    cd.Pr[t] = float(FT(_per_tracer(c.Pr, self.tracer_names, t)))
            elif isinstance(c, AnisotropicMinimumDissipation):
                cd.kind, cd.Cnu = _abi.OB_CLOSURE_AMD, float(FT(c.Cnu))
                cd.vertically_implicit = int(c.vertically_implicit)
                cd.amd_has_cb, cd.cb = (0, 0.0) if c.Cb is None else (1, float(FT(c.Cb)))
                for t in range(nt):
                    cd.Ckappa[t] = float(FT(_per_tracer(c.Ckappa, self.tracer_names, t)))
            else:
                raise _abi.OceanB200Error(-3, "closure %r is outside the B200 hot path" % (c,))
            if getattr(c, "vertically_implicit", False) and g.topo[2] != _abi.OB_BOUNDED:
                raise ValueError("VerticallyImplicitTimeDiscretization can only be specified on grids that are Bounded in the z-direction.")
        if buoyancy is None:
            d.buoyancy_kind = _abi.OB_BUOYANCY_NONE
        elif isinstance(buoyancy, BuoyancyTracer):
            if "b" not in self.tracer_names:
                raise ValueError("BuoyancyTracer() requires a tracer named b")
            d.buoyancy_kind, d.buoyancy_tracer = _abi.OB_BUOYANCY_TRACER, self.tracer_names.index("b")
        elif isinstance(buoyancy, SeawaterBuoyancy):
            if "T" not in self.tracer_names or "S" not in self.tracer_names:
                raise ValueError("SeawaterBuoyancy() requires tracers T and S")
            d.buoyancy_kind = _abi.OB_BUOYANCY_LINEAR_SEAWATER
            d.temperature_tracer, d.salinity_tracer = self.tracer_names.index("T"), self.tracer_names.index("S")
            d.g = float(FT(buoyancy.g))
            d.thermal_expansion = float(FT(buoyancy.equation_of_state.thermal_expansion))
            d.haline_contraction = float(FT(buoyancy.equation_of_state.haline_contraction))
        else:
            raise _abi.OceanB200Error(-3, "buoyancy %r is outside the B200 hot path" % (buoyancy,))
        if coriolis is not None:
            if not isinstance(coriolis, FPlane):
                raise _abi.OceanB200Error(-3, "only FPlane Coriolis is on the B200 hot path")
            d.has_coriolis, d.f = 1, float(FT(coriolis.f))
        d.n_tracers = nt
        d.stepper = _abi.OB_RK3 if ts == "RungeKutta3" else _abi.OB_AB2
        d.chi = float(self.chi)
        d.has_hydrostatic_pressure = int(has_phy)
        d.bcs_u = bc_desc(self.velocities["u"].boundary_conditions)
        d.bcs_v = bc_desc(self.velocities["v"].boundary_conditions)
        d.bcs_w = bc_desc(self.velocities["w"].boundary_conditions)
        d.bcs_p = bc_desc(self.pressures["pNHS"].boundary_conditions)
        if has_phy:
            d.bcs_phy = bc_desc(self.pressures["pHY"].boundary_conditions)
        for t, n in enumerate(self.tracer_names):
            d.bcs_tracer[t] = bc_desc(self.tracers[n].boundary_conditions)
        for m, cf in enumerate(self.closure_fields):
            if "nue" in cf:
                d.bcs_nue[m] = bc_desc(cf["nue"].boundary_conditions)
            for t, f in enumerate(cf.get("kappae", [])):
                d.bcs_kappae[m][t] = bc_desc(f.boundary_conditions)
        self.desc = d
        h = C.c_void_p()
        _abi.call("ob_model_create", arch.ctx, C.byref(d), C.byref(h))
        self.handle = h
        self._bind_all()
        self._upload_bc_arrays()
        self.update_state()

    # ---------------------------------------------------------------------------------------------------------
    def _bind(self, fid, f):
        _abi.call("ob_model_bind_field", self.handle, fid, f.data)

    def _bind_all(self):
        A = _abi
        self._bind(A.OB_FIELD_U, self.velocities["u"]); self._bind(A.OB_FIELD_V, self.velocities["v"]); self._bind(A.OB_FIELD_W, self.velocities["w"])
        self._bind(A.OB_FIELD_PNHS, self.pressures["pNHS"])
        if "pHY" in self.pressures:
            self._bind(A.OB_FIELD_PHY, self.pressures["pHY"])
        for t, n in enumerate(self.tracer_names):
            self._bind(A.OB_FIELD_TRACER0 + t, self.tracers[n])
        for n in range(3 + len(self.tracer_names)):
            self._bind(A.OB_FIELD_GN0 + n, self.Gn[n]); self._bind(A.OB_FIELD_GM0 + n, self.Gm[n])
        for m, cf in enumerate(self.closure_fields):
            if "nue" in cf:
                self._bind(A.OB_FIELD_NUE0 + m, cf["nue"])
            for t, f in enumerate(cf.get("kappae", [])):
                self._bind(A.OB_FIELD_KAPPAE0 + m * A.OB_MAX_TRACERS + t, f)

    def _upload_bc_arrays(self):
        """array-valued boundary conditions: device copies registered with ob_model_set_bc_array (kept alive here)"""
        from .fields import SIDES
        A = _abi
        self._bc_arrays = []
        ids = {"u": A.OB_FIELD_U, "v": A.OB_FIELD_V, "w": A.OB_FIELD_W}
        ids.update({n: A.OB_FIELD_TRACER0 + t for t, n in enumerate(self.tracer_names)})
        for name, f in self.prognostic_fields.items():
            for s, side in enumerate(SIDES):
                bc = f.boundary_conditions[side]
                if bc is None or not isinstance(bc.value, np.ndarray):
                    continue
                d = s // 2
                da, db = (1 if d == 0 else 0), (1 if d == 2 else 2)
                want = (f.n[db], f.n[da])
                arr = np.ascontiguousarray(bc.value, dtype=self.grid.FT)
                if arr.shape != want:
                    raise ValueError("%s boundary condition of %s: array shape %r, expected %r (n_slow, n_fast)" % (side, name, arr.shape, want))
                p = C.c_void_p()
                _abi.call("ob_malloc", self.architecture.ctx, arr.nbytes, C.byref(p))
                _abi.call("ob_memcpy_h2d", self.architecture.ctx, p, arr.ctypes.data_as(C.c_void_p), arr.nbytes)
                _abi.call("ob_sync", self.architecture.ctx)
                _abi.call("ob_model_set_bc_array", self.handle, ids[name], s, p)
                self._bc_arrays.append(p)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _abi.lib().ob_model_destroy(self.handle)
                self.handle = None
            for p in getattr(self, "_bc_arrays", []):
                _abi.lib().ob_free(self.architecture.ctx, p)
            self._bc_arrays = []
        except Exception:
            pass

    @property
    def prognostic_fields(self):
        out = dict(self.velocities)
        out.update(self.tracers)
        return out

    def field_id(self, name):
        if name in ("u", "v", "w"):
            return "uvw".index(name)
        if name == "pNHS":
            return _abi.OB_FIELD_PNHS
        if name == "pHY":
            return _abi.OB_FIELD_PHY
        return _abi.OB_FIELD_TRACER0 + self.tracer_names.index(name)

    # fine-grained entry points (what the Julia shim's method overrides call) ----------------------------------
    def fill_halo_regions(self, name, fill_normal_flow_bcs=True):
        _abi.call("ob_fill_halo", self.handle, self.field_id(name), int(fill_normal_flow_bcs))

    def update_state(self):
        _abi.call("ob_update_state", self.handle)

    def compute_tendencies(self):
        _abi.call("ob_compute_tendencies", self.handle)

    def rk3_substep(self, dt, gamma, zeta):
        _abi.call("ob_rk3_substep", self.handle, float(dt), float(gamma), 0.0 if zeta is None else float(zeta), int(zeta is not None))

    def ab2_step(self, dt, chi):
        _abi.call("ob_ab2_step", self.handle, float(dt), float(chi))

    def cache_previous_tendencies(self):
        _abi.call("ob_cache_tendencies", self.handle)

    def compute_pressure_correction(self, dtau):
        _abi.call("ob_compute_pressure_correction", self.handle, float(dtau))

    def make_pressure_correction(self, dtau):
        _abi.call("ob_make_pressure_correction", self.handle, float(dtau))

    def cell_advection_timescale(self):
        tau = C.c_double(0)
        _abi.call("ob_cell_advection_timescale", self.handle, C.byref(tau))
        from .distributed import all_reduce_scalar
        return all_reduce_scalar(self.architecture, tau.value, "min")

    def launch_count(self):
        n = C.c_int64(0)
        _abi.call("ob_launch_count", self.handle, C.byref(n))
        return n.value

    def set_option(self, option, value):
        _abi.call("ob_model_set_option", self.handle, int(option), int(value))

    def synchronize(self):
        self.architecture.synchronize()


def update_state(model):
    model.update_state()


def set(model, enforce_incompressibility=True, **kw):
    """set!(model; u=..., T=...) (set_nonhydrostatic_model.jl:39-74): per-field set + halo fill, update_state!,
    then a pressure projection with Δt = 1 and another update_state!."""
    fields = model.prognostic_fields
    for name, val in kw.items():
        if name not in fields:
            raise ValueError("name %s not found in model.velocities or model.tracers." % name)
        fields[name].set(val)
        model.fill_halo_regions(name)
    model.update_state()
    if enforce_incompressibility:
        model.compute_pressure_correction(1.0)
        model.make_pressure_correction(1.0)
        model.update_state()


def time_step(model, dt, euler=False, callbacks=()):
    """time_step!(model, Δt).  With no callbacks the whole step is ONE C-ABI call (ob_time_step_rk3/ab2); with
    `callbacks` (called after every update_state!, like TendencyCallsite/UpdateStateCallsite users) the stages are
    driven from the host through the fine-grained entry points."""
    clk = model.clock
    FT = model.grid.FT
    first = int(clk.iteration == 0)
    if model.timestepper == "RungeKutta3":
        if first:
            clk.last_dt = dt
        if not callbacks:
            _abi.call("ob_time_step_rk3", model.handle, float(dt), first)
        else:
            g1, g2, g3 = FT(8.0 / 15.0), FT(5.0 / 12.0), FT(3.0 / 4.0)
            z2, z3 = FT(-17.0 / 60.0), FT(-5.0 / 12.0)
            if first:
                model.update_state()
            for gam, zet in ((g1, None), (g2, z2), (g3, z3)):
                model.rk3_substep(dt, gam, zet)
                model.cache_previous_tendencies()
                model.update_state()
                for cb in callbacks:
                    cb(model)
        # clock: tⁿ⁺¹ computed a priori (runge_kutta_3.jl:121-122, clock.jl:147-173)
        clk.time = clk.time + dt
        clk.last_stage_dt = dt * float(FT(3.0 / 4.0) + FT(-5.0 / 12.0))
    else:
        euler = bool(euler) or (dt != clk.last_dt)
        if not callbacks:
            _abi.call("ob_time_step_ab2", model.handle, float(dt), int(euler), first)
        else:
            if first:
                model.update_state()
            model.ab2_step(dt, -0.5 if euler else float(model.chi))
            model.cache_previous_tendencies()
            model.update_state()
            for cb in callbacks:
                cb(model)
        clk.time = clk.time + dt
        clk.last_stage_dt = dt
    clk.iteration += 1
    clk.last_dt = dt
