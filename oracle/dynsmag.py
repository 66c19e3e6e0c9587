"""TEST INFRASTRUCTURE (oracle): DynamicSmagorinsky with a directionally averaged coefficient, restated with numpy.

Reference: src/TurbulenceClosures/turbulence_closure_implementations/Smagorinskys/
  dynamic_coefficient.jl:245-351      square_smagorinsky_coefficient, _compute_Σ!, _compute_Σ̄!, _compute_LM_MM!, LM_and_MM,
                                      compute_coefficient_fields!(::DirectionallyAveragedDynamicSmagorinsky), allocate_coefficient_fields
  scale_invariant_operators.jl:10-188 ΣᵢⱼΣᵢⱼᶜᶜᶜ, filter, filtered gradients and strains, Σ̄ᵢⱼΣ̄ᵢⱼᶜᶜᶜ, ⟨ΣΣᵢⱼ⟩, Σ̄Σ̄ᵢⱼ, Mᵢⱼ, Lᵢⱼ
  smagorinsky.jl:90-104               νₑ = cˢ² Δᶠ² √(2 Σ²)
(Bou-Zeid, Meneveau & Parlange 2005, scale-invariant form: ᾱ² = 4, β = 1.)  LagrangianAveraging is not restated.

Every quantity is a `GF`: values over the logical window [1 - pad, N + pad]^3 of the cell-centred index space; a one-cell shift
costs one cell of pad, binary operations trim to the smaller pad.  Raw fields are embedded in a window of pad = PAD >= H that is
NaN outside the parent array: a chain of shifts that really reached beyond the halo would put NaNs into the interior, which is
asserted not to happen (the deepest chains reach two cells: H >= 2 is required, as for the reference's own kernels).  Flat
directions are not supported.
"""
import numpy as np

ALPHA2_BETA = 4 * 1   # ᾱ² * β (scale_invariant_operators.jl:143-144)
PAD = 8


class GF:
    def __init__(self, a, pad):
        self.a, self.pad = a, pad

    def trim(self, q):
        d = self.pad - q
        return self.a if d == 0 else self.a[d:-d, d:-d, d:-d]

    def sh(self, di=0, dj=0, dk=0):
        """value at (i + di, j + dj, k + dk), shifts in {-1, 0, 1}"""
        assert self.pad >= 1 and max(abs(di), abs(dj), abs(dk)) <= 1
        nk, nj, ni = (m - 2 for m in self.a.shape)
        return GF(self.a[1 + dk:1 + dk + nk, 1 + dj:1 + dj + nj, 1 + di:1 + di + ni], self.pad - 1)

    def _bin(self, o, f):
        if not isinstance(o, GF):
            return GF(f(self.a, o), self.pad)
        q = min(self.pad, o.pad)
        return GF(f(self.trim(q), o.trim(q)), q)

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return GF(o * self.a, self.pad)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def sq(self): return GF(self.a * self.a, self.pad)   # x^2 (Base.literal_pow: x*x)
    def sqrt(self): return GF(np.sqrt(self.a), self.pad)
    def interior(self):
        a = self.trim(0)
        assert not np.isnan(a).any(), "a stencil reached beyond the halo"
        return a


def _field_gf(f):
    """window [1-H, N+H] of a field in the cell-centred index space (Face fields have one more point: it is not needed)"""
    g = f.grid
    H, N = g.H, g.N
    a = np.full((N[2] + 2 * PAD, N[1] + 2 * PAD, N[0] + 2 * PAD), np.nan, dtype=g.ft)
    a[PAD - H[2]:PAD + N[2] + H[2], PAD - H[1]:PAD + N[1] + H[1], PAD - H[0]:PAD + N[0] + H[0]] = \
        f.view(1 - H[0], N[0] + H[0], 1 - H[1], N[1] + H[1], 1 - H[2], N[2] + H[2])
    return GF(a, PAD)


def _metric_gf(grid, which):
    """1/Δzᶜ(k), 1/Δzᶠ(k), Δzᶜ(k) as full GFs (regular x, y enter as scalars)"""
    ft, H, N = grid.ft, grid.H, grid.N
    ks = np.arange(1 - H[2], N[2] + H[2] + 1)
    v = {"rdzc": lambda: ft(1) / grid.dC(2, ks).astype(ft), "rdzf": lambda: ft(1) / grid.dF(2, ks).astype(ft),
         "dzc": lambda: grid.dC(2, ks).astype(ft)}[which]()
    col = np.full(N[2] + 2 * PAD, np.nan, dtype=ft)
    col[PAD - H[2]:PAD + N[2] + H[2]] = v
    shape = (N[2] + 2 * PAD, N[1] + 2 * PAD, N[0] + 2 * PAD)
    return GF(np.broadcast_to(col[:, None, None], shape).astype(ft), PAD)


class _Ops:
    def __init__(self, grid):
        ft = grid.ft
        self.ft = ft
        self.h = ft(0.5)
        self.rdx = ft(1) / ft(grid.dC(0, np.array([1]))[0])
        self.rdy = ft(1) / ft(grid.dC(1, np.array([1]))[0])
        self.rdzc, self.rdzf, self.dzc = (_metric_gf(grid, w) for w in ("rdzc", "rdzf", "dzc"))
        self.dx = ft(grid.dC(0, np.array([1]))[0])
        self.dy = ft(grid.dC(1, np.array([1]))[0])

    # interpolation to centres / faces (interpolation_operators.jl:8-28)
    def Ixc(self, f): return self.h * (f + f.sh(di=1))
    def Iyc(self, f): return self.h * (f + f.sh(dj=1))
    def Izc(self, f): return self.h * (f + f.sh(dk=1))
    def Ixyc(self, f): return self.Iyc(self.Ixc(f))   # ℑxyᶜᶜᵃ = ℑyᵃᶜᵃ(ℑxᶜᵃᵃ)
    def Ixzc(self, f): return self.Izc(self.Ixc(f))   # ℑxzᶜᵃᶜ = ℑzᵃᵃᶜ(ℑxᶜᵃᵃ)
    def Iyzc(self, f): return self.Izc(self.Iyc(f))   # ℑyzᵃᶜᶜ = ℑzᵃᵃᶜ(ℑyᵃᶜᵃ)

    # the six strain components from (possibly filtered) velocities (strain / scale_invariant_operators.jl:60-110)
    def strains(self, u, v, w):
        s11 = (u.sh(di=1) - u) * self.rdx                      # ∂xᶜᶜᶜ u
        s22 = (v.sh(dj=1) - v) * self.rdy
        s33 = (w.sh(dk=1) - w) * self.rdzc
        dyu = (u - u.sh(dj=-1)) * self.rdy                     # ∂yᶠᶠᶜ u
        dxv = (v - v.sh(di=-1)) * self.rdx                     # ∂xᶠᶠᶜ v
        dzu = (u - u.sh(dk=-1)) * self.rdzf                    # ∂zᶠᶜᶠ u
        dxw = (w - w.sh(di=-1)) * self.rdx                     # ∂xᶠᶜᶠ w
        dzv = (v - v.sh(dk=-1)) * self.rdzf                    # ∂zᶜᶠᶠ v
        dyw = (w - w.sh(dj=-1)) * self.rdy                     # ∂yᶜᶠᶠ w
        s12 = self.h * (dyu + dxv)
        s13 = self.h * (dzu + dxw)
        s23 = self.h * (dzv + dyw)
        return s11, s22, s33, s12, s13, s23

    def double_dot(self, st):
        """ΣᵢⱼΣᵢⱼᶜᶜᶜ (scale_invariant_operators.jl:10-13 / :112-116)"""
        s11, s22, s33, s12, s13, s23 = st
        tr = s11.sq() + s22.sq() + s33.sq()
        return tr + 2 * self.Ixyc(s12.sq()) + 2 * self.Ixzc(s13.sq()) + 2 * self.Iyzc(s23.sq())

    def filt(self, f):
        """filter (scale_invariant_operators.jl:47-57)"""
        s = 6 * f + f.sh(di=1) + f.sh(di=-1) + f.sh(dj=1) + f.sh(dj=-1) + f.sh(dk=1) + f.sh(dk=-1)
        return s / self.ft(12)


def _center_field_like(model, name):
    from .model import Field
    return Field(model.grid, "ccc", None, name)


def compute_dynamic_smagorinsky(model, m):
    """compute_closure_fields!(closure_fields, ::DirectionallyAveragedDynamicSmagorinsky, model): coefficient fields, then νₑ"""
    from .model import fill_halo_regions, FLAT
    c = model.closures[m]
    g = model.grid
    ft = g.ft
    if any(t == FLAT for t in g.topo):
        raise ValueError("DynamicSmagorinsky: Flat directions are not supported")
    if min(g.H) < 2:
        raise ValueError("DynamicSmagorinsky needs a halo of at least 2")
    O = _Ops(g)
    u, v, w = _field_gf(model.u), _field_gf(model.v), _field_gf(model.w)
    st = O.strains(u, v, w)
    ub, vb, wb = O.filt(u), O.filt(v), O.filt(w)
    stb = O.strains(ub, vb, wb)
    # _compute_Σ!, _compute_Σ̄! over :xyz, then their halos (dynamic_coefficient.jl:312-327)
    S2 = O.double_dot(st)
    cf = model.dynamic_fields.setdefault(m, {k: _center_field_like(model, k) for k in ("Sigma", "Sigmabar", "LM", "MM")})
    cf["Sigma"].interior[...] = np.sqrt(S2.interior())
    cf["Sigmabar"].interior[...] = np.sqrt(O.double_dot(stb).interior())
    fill_halo_regions(cf["Sigma"]); fill_halo_regions(cf["Sigmabar"])
    Sg, Sb = _field_gf(cf["Sigma"]), _field_gf(cf["Sigmabar"])
    s11, s22, s33, s12, s13, s23 = st
    b11, b22, b33, b12, b13, b23 = stb
    # ⟨ΣΣᵢⱼ⟩ and Σ̄Σ̄ᵢⱼ at ccc (scale_invariant_operators.jl:120-141)
    SS = [Sg * s11, Sg * s22, Sg * s33, Sg * O.Ixyc(s12), Sg * O.Ixzc(s13), Sg * O.Iyzc(s23)]
    fSS = [O.filt(x) for x in SS]
    BB = [Sb * b11, Sb * b22, Sb * b33, Sb * O.Ixyc(b12), Sb * O.Ixzc(b13), Sb * O.Iyzc(b23)]
    # Δᶠ = ∛volume, Mᵢⱼ = 2 Δᶠ² (⟨ΣΣᵢⱼ⟩ - ᾱ² β Σ̄Σ̄ᵢⱼ)  (:145-153)
    Df = GF(np.cbrt(((O.dx * O.dy) * O.dzc.a).astype(ft)), O.dzc.pad)
    twoD2 = 2 * Df.sq()
    M = [twoD2 * (a - ft(ALPHA2_BETA) * b) for a, b in zip(fSS, BB)]
    # Lᵢⱼ = filter(uᵢuⱼ) - ūᵢūⱼ at ccc (:155-188)
    uc, vc, wc = O.Ixc(u), O.Iyc(v), O.Izc(w)
    ubc, vbc, wbc = O.Ixc(ub), O.Iyc(vb), O.Izc(wb)
    L = [O.filt(O.Ixc(u.sq())) - O.Ixc(ub.sq()), O.filt(O.Iyc(v.sq())) - O.Iyc(vb.sq()), O.filt(O.Izc(w.sq())) - O.Izc(wb.sq()),
         O.filt(uc * vc) - ubc * vbc, O.filt(uc * wc) - ubc * wbc, O.filt(vc * wc) - vbc * wbc]
    LM = L[0] * M[0] + L[1] * M[1] + L[2] * M[2] + (2 * L[3]) * M[3] + (2 * L[4]) * M[4] + (2 * L[5]) * M[5]
    MM = M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + (2 * M[3]) * M[3] + (2 * M[4]) * M[4] + (2 * M[5]) * M[5]
    cf["LM"].interior[...] = LM.interior()
    cf["MM"].interior[...] = MM.interior()
    # 𝒥ᴸᴹ = Average(LM, dims), 𝒥ᴹᴹ = Average(MM, dims) over the interior (:343-344); numpy axes are (k, j, i)
    axes = tuple(2 - (d - 1) for d in sorted(c.dynamic["averaging"]))
    JLM = cf["LM"].interior.astype(np.float64).mean(axis=axes, keepdims=True).astype(ft)
    JMM = cf["MM"].interior.astype(np.float64).mean(axis=axes, keepdims=True).astype(ft)
    cf["JLM"], cf["JMM"] = JLM, JMM
    # cˢ² = max(𝒥ᴸᴹ, 𝒥ᴸᴹ_min) / 𝒥ᴹᴹ * (𝒥ᴹᴹ > 0)  (:245-256); νₑ = cˢ² Δᶠ² √(2Σ²) (smagorinsky.jl:90-104)
    num = np.maximum(JLM, ft(c.dynamic["minimum_numerator"]))
    with np.errstate(divide="ignore", invalid="ignore"):
        cs2 = np.where(JMM > 0, num / JMM, ft(0)).astype(ft)
    D3 = ((O.dx * O.dy) * O.dzc.interior()).astype(ft)
    Dfv = np.cbrt(D3)
    model.nue[m].interior[...] = (cs2 * (Dfv * Dfv) * np.sqrt(2 * S2.interior())).astype(ft)
    return cs2
