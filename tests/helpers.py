"""Shared test helpers: build the SAME configuration as a CPU-oracle model (oracle/model.py) and as a B200 model
(ocean_b200, through the C ABI), seed both with identical initial conditions, compare fields."""
import numpy as np


def _oracle():
    """the CPU oracle is imported lazily: bench.py's GPU arm uses Config without ever touching oracle/"""
    from oracle import model as M
    return M

TOPO_CLS = {"P": "Periodic", "B": "Bounded", "F": "Flat"}


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.sqrt(np.sum((a - b) ** 2))
    n = np.sqrt(np.sum(b ** 2))
    return d / n if n > 0 else d


def stretched_faces(Nz, Lz, sigma=1.1):
    """z(k) = -Lz (1 - tanh(σ (k-1)/Nz) / tanh σ)  (rectilinear_grid.jl docstring :218-228)"""
    k = np.arange(1, Nz + 2)
    return -Lz * (1 - np.tanh(sigma * (k - 1) / Nz) / np.tanh(sigma))


class Config:
    def __init__(self, size, extent, topology="PPB", halo=(3, 3, 3), ft=np.float64, advection=("weno", 5), closure=(),
                 buoyancy=None, coriolis_f=None, tracers=(), timestepper="rk3", bcs=None, weno_division=None):
        self.size, self.extent, self.topology, self.halo, self.ft = size, extent, topology, halo, ft
        self.advection, self.closure, self.buoyancy, self.coriolis_f = advection, tuple(closure), buoyancy, coriolis_f
        self.tracers, self.timestepper, self.bcs = tuple(tracers), timestepper, bcs or {}
        self.weno_division = weno_division

    # ---- oracle ---------------------------------------------------------------------------------------------
    def oracle_grid(self):
        return _oracle().Grid(self.size, self.extent, topology=tuple(self.topology), halo=self.halo, ft=self.ft)

    def oracle_model(self):
        M = _oracle()
        cl = []
        for c in self.closure:
            if c[0] == "scalar":
                cl.append(M.ScalarDiffusivity(nu=c[1], kappa=c[2]))
            elif c[0] == "vi_scalar":
                cl.append(M.ScalarDiffusivity(nu=c[1], kappa=c[2], vertically_implicit=True))
            elif c[0] in ("smag", "vi_smag"):
                cl.append(M.Smagorinsky(coefficient=c[1], Pr=c[2], vertically_implicit=c[0] == "vi_smag"))
            elif c[0] in ("lilly", "vi_lilly"):
                cl.append(M.SmagorinskyLilly(C=c[1], Cb=c[2], Pr=c[3], vertically_implicit=c[0] == "vi_lilly"))
            elif c[0] in ("amd", "vi_amd"):
                cl.append(M.AnisotropicMinimumDissipation(Cb=c[1] if len(c) > 1 else None, vertically_implicit=c[0] == "vi_amd"))
            elif c[0] in ("dynsmag", "vi_dynsmag"):   # ("dynsmag", averaging dims, Pr)
                cl.append(M.DynamicSmagorinsky(averaging=c[1], Pr=c[2], vertically_implicit=c[0] == "vi_dynsmag"))
        bcs = {}
        for name, sides in self.bcs.items():
            bcs[name] = {s: (k.lower(), v) for s, (k, v) in sides.items()}
        return M.Model(self.oracle_grid(), advection=self.advection, closure=cl or None, buoyancy=self.buoyancy,
                       coriolis_f=self.coriolis_f, tracers=self.tracers, timestepper=self.timestepper, boundary_conditions=bcs)

    # ---- B200 -----------------------------------------------------------------------------------------------
    def b200_grid(self, arch):
        import ocean_b200 as ob
        topo = tuple(getattr(ob, TOPO_CLS[t]) for t in self.topology)
        nonflat = [d for d in range(3) if self.topology[d] != "F"]
        kw = {}
        for d in nonflat:
            e = self.extent[d]
            kw["xyz"[d]] = tuple(e) if isinstance(e, tuple) else np.asarray(e)
        return ob.RectilinearGrid(arch, self.ft, size=tuple(self.size[d] for d in nonflat),
                                  halo=tuple(self.halo[d] for d in nonflat), topology=topo, **kw)

    def b200_model(self, arch):
        import ocean_b200 as ob
        adv = self.advection
        advection = ob.WENO(order=adv[1], weight_computation=self.weno_division) if adv[0] == "weno" else ob.Centered(order=adv[1])
        cl = []
        for c in self.closure:
            if c[0] == "scalar":
                cl.append(ob.ScalarDiffusivity(nu=c[1], kappa=c[2]))
            elif c[0] == "vi_scalar":
                cl.append(ob.ScalarDiffusivity(ob.VerticallyImplicitTimeDiscretization(), nu=c[1], kappa=c[2]))
            elif c[0] in ("smag", "vi_smag"):
                td = (ob.VerticallyImplicitTimeDiscretization(),) if c[0] == "vi_smag" else ()
                cl.append(ob.Smagorinsky(*td, coefficient=c[1], Pr=c[2]))
            elif c[0] in ("lilly", "vi_lilly"):
                td = (ob.VerticallyImplicitTimeDiscretization(),) if c[0] == "vi_lilly" else ()
                cl.append(ob.SmagorinskyLilly(*td, C=c[1], Cb=c[2], Pr=c[3]))
            elif c[0] in ("amd", "vi_amd"):
                td = (ob.VerticallyImplicitTimeDiscretization(),) if c[0] == "vi_amd" else ()
                cl.append(ob.AnisotropicMinimumDissipation(*td, Cb=c[1] if len(c) > 1 else None))
            elif c[0] in ("dynsmag", "vi_dynsmag"):
                td = (ob.VerticallyImplicitTimeDiscretization(),) if c[0] == "vi_dynsmag" else ()
                cl.append(ob.DynamicSmagorinsky(*td, averaging=c[1], Pr=c[2]))
        b = self.buoyancy
        if b is None:
            buoy = None
        elif b[0] == "tracer":
            buoy = ob.BuoyancyTracer()
        else:
            buoy = ob.SeawaterBuoyancy(gravitational_acceleration=b[1],
                                       equation_of_state=ob.LinearEquationOfState(thermal_expansion=b[2], haline_contraction=b[3]))
        mk = {"Flux": ob.FluxBoundaryCondition, "Value": ob.ValueBoundaryCondition, "Gradient": ob.GradientBoundaryCondition}
        bcs = {name: ob.FieldBoundaryConditions(**{s: mk[k](v) for s, (k, v) in sides.items()}) for name, sides in self.bcs.items()}
        cor = None if self.coriolis_f is None else ob.FPlane(f=self.coriolis_f)
        closure = None if not cl else (cl[0] if len(cl) == 1 else tuple(cl))
        ts = "RungeKutta3" if self.timestepper == "rk3" else "QuasiAdamsBashforth2"
        return ob.NonhydrostaticModel(self.b200_grid(arch), advection=advection, closure=closure, buoyancy=buoy, coriolis=cor,
                                      tracers=self.tracers, timestepper=ts, boundary_conditions=bcs)

    # ---- initial conditions ---------------------------------------------------------------------------------
    def initial_conditions(self, seed, amp=0.1, tracer_amp=1e-3, tracer_mean=None):
        """interior arrays (numpy (nz, ny, nx)) for every prognostic field; identical for both models"""
        rng = np.random.default_rng(seed)
        N = self.size
        bounded = [t == "B" for t in self.topology]

        def shape(loc):  # interior shape (nz, ny, nx): +1 for a Face location along a Bounded direction
            return tuple(N[d] + (1 if (loc[d] == "f" and bounded[d]) else 0) for d in (2, 1, 0))

        out = {}
        for name, loc in (("u", "fcc"), ("v", "cfc"), ("w", "ccf")):
            d = "uvw".index(name)
            a = amp * rng.uniform(-1, 1, shape(loc))
            if any(c[0] in ("dynsmag", "vi_dynsmag") for c in self.closure):
                # a resolved large-scale flow under the noise: the dynamic procedure returns c_s = 0 for pure noise (<LM> < 0)
                kk, jj, ii = np.meshgrid(*(np.arange(n) / max(n - 1, 1) for n in shape(loc)), indexing="ij")
                ph = 2 * np.pi
                a = a + 5 * amp * [np.sin(ph * ii) * np.cos(ph * jj) * np.cos(ph * kk) + 0.3 * np.sin(2 * ph * ii + 1) * np.cos(3 * ph * jj) * np.cos(2 * ph * kk),
                                   -np.cos(ph * ii) * np.sin(ph * jj) * np.cos(ph * kk) + 0.3 * np.cos(2 * ph * ii) * np.sin(3 * ph * jj + 2) * np.cos(ph * kk),
                                   0.4 * np.cos(2 * ph * ii) * np.cos(ph * jj) * np.sin(2 * ph * kk)][d]
            if self.topology[d] == "F":
                a[...] = 0  # no flow in a Flat direction
            out[name] = a.astype(self.ft)
        if self.topology[2] == "F":
            zc = 0.0
        else:
            e = self.extent[2]
            faces = np.linspace(e[0], e[1], N[2] + 1) if isinstance(e, tuple) else np.asarray(e, dtype=np.float64)
            zc = (0.5 * (faces[1:] + faces[:-1])).astype(self.ft)[:, None, None]
        for t in self.tracers:
            mean = (tracer_mean or {}).get(t, 0.0)
            base = {"b": 1.0 * zc, "T": 20 + 0.005 * zc, "S": 35.0 + 0 * zc}.get(t, 0 * zc) + mean
            out[t] = (base + tracer_amp * rng.uniform(-1, 1, shape("ccc"))).astype(self.ft)
        return out


def pair(cfg, arch, seed=1, **ickw):
    """(oracle model, b200 model) with identical initial conditions set through `set!`"""
    import ocean_b200 as ob
    ic = cfg.initial_conditions(seed, **ickw)
    om = cfg.oracle_model()
    om.set(**ic)
    bm = cfg.b200_model(arch)
    ob.set(bm, **ic)
    return om, bm


def oracle_fields(om):
    out = {"u": om.u.data, "v": om.v.data, "w": om.w.data, "pNHS": om.pNHS.data}
    for n, f in zip(om.tracer_names, om.tracers):
        out[n] = f.data
    return out


def b200_fields(bm):
    out = {k: f.parent() for k, f in bm.velocities.items()}
    out["pNHS"] = bm.pressures["pNHS"].parent()
    for n, f in bm.tracers.items():
        out[n] = f.parent()
    return out


def interior_of(om_field, arr):
    """interior view of a parent-shaped array using the oracle field's geometry"""
    H = om_field.grid.H
    n = om_field.n
    return arr[H[2]:H[2] + n[2], H[1]:H[1] + n[1], H[0]:H[0] + n[0]]
