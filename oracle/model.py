"""CPU oracle of the Oceananigans.jl NonhydrostaticModel time step (TEST INFRASTRUCTURE ONLY).

numpy orchestration + the C pointwise kernels of oracle/csrc/oracle.c, following the reference
call stack:

  time_step!            src/TimeSteppers/runge_kutta_3.jl:103-168, quasi_adams_bashforth_2.jl:90-126
  rk3_substep!          src/Models/NonhydrostaticModels/nonhydrostatic_rk3_substep.jl:31-63
  pressure correction   .../pressure_correction.jl:6-106, solve_for_pressure.jl:12-126
  update_state!         .../update_nonhydrostatic_model_state.jl:22-83
  fill_halo_regions!    src/BoundaryConditions/fill_halo_regions*.jl, boundary_condition_ordering.jl
  Poisson solvers       src/Solvers/fft_based_poisson_solver.jl:94-124,
                        fourier_tridiagonal_poisson_solver.jl:199-260, batched_tridiagonal_solver.jl:211-243
                        (FFTW is replaced by scipy.fft / pocketfft: same transforms, same normalisation)

Arrays are numpy with shape (Pz, Py, Px), i.e. x fastest, identical in memory to the reference's
column-major parent arrays of size (Px, Py, Pz).

PARITY: unpinned against reference outputs (see oracle.c header); pinned by known-answer tests.
"""
import ctypes as C
import os

import numpy as np
import scipy.fft as sfft

from . import build as _build
from . import coefficients as coef

MAXTR, MAXCL, MAXBUF = 8, 4, 6
PERIODIC, BOUNDED, FLAT = 0, 1, 2
TOPO = {"P": PERIODIC, "B": BOUNDED, "F": FLAT, "Periodic": PERIODIC, "Bounded": BOUNDED, "Flat": FLAT}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
    return _lib


def set_num_threads(n):
    """force the OpenMP team size of the C kernels (whatever OMP_NUM_THREADS says) and return the size in effect"""
    lib().orc_set_num_threads(int(n))
    return int(lib().orc_get_max_threads())


def _mkstructs(ft):
    cft = C.c_double if ft == np.float64 else C.c_longdouble if ft == np.longdouble else C.c_float
    P = C.POINTER(cft)

    class OField(C.Structure):
        _fields_ = [("p", P), ("P", C.c_int * 3), ("o", C.c_int * 3)]

    class OParams(C.Structure):
        _fields_ = [
            ("N", C.c_int * 3), ("H", C.c_int * 3), ("topo", C.c_int * 3),
            ("df", P * 3), ("dc", P * 3), ("dfo", C.c_int * 3), ("dco", C.c_int * 3),
            ("adv_kind", C.c_int), ("adv_buffer", C.c_int),
            ("weno_coeff", P), ("weno_beta", P), ("weno_cstar", P), ("cen_coeff", P), ("weno_eps", cft),
            ("nclosures", C.c_int), ("closure_kind", C.c_int * MAXCL),
            ("nu", cft * MAXCL), ("kappa", (cft * MAXTR) * MAXCL), ("Pr", (cft * MAXTR) * MAXCL),
            ("nue", OField * MAXCL), ("kappae", (OField * MAXTR) * MAXCL),
            ("cs", cft * MAXCL), ("cb", cft * MAXCL), ("lilly", C.c_int * MAXCL),
            ("Cnu", cft * MAXCL), ("Ckappa", (cft * MAXTR) * MAXCL), ("amd_has_cb", C.c_int * MAXCL),
            ("buoy_kind", C.c_int), ("ib", C.c_int), ("iT", C.c_int), ("iS", C.c_int),
            ("g", cft), ("alpha", cft), ("beta", cft),
            ("has_coriolis", C.c_int), ("fcor", cft),
            ("ntracers", C.c_int),
            ("u", OField), ("v", OField), ("w", OField), ("c", OField * MAXTR),
            ("has_pHY", C.c_int), ("pHY", OField),
            ("Gu", OField), ("Gv", OField), ("Gw", OField), ("Gc", OField * MAXTR),
            ("closure_vi", C.c_int * MAXCL),
        ]

    return cft, P, OField, OParams


_STRUCTS = {}


def structs(ft):
    ft = np.dtype(ft).type
    if ft not in _STRUCTS:
        _STRUCTS[ft] = _mkstructs(ft)
    return _STRUCTS[ft]


# ---------------------------------------------------------------------------------------------
# Grid (src/Grids/rectilinear_grid.jl:266-294, grid_generation.jl:34-156)
# ---------------------------------------------------------------------------------------------
class Grid:
    """RectilinearGrid restatement.  `extent` entries are (lo, hi) tuples for regular directions or a
    1-D array of N+1 face positions for a stretched direction; Flat directions are `None`."""

    def __init__(self, size, extent, topology=("P", "P", "B"), halo=(3, 3, 3), ft=np.float64):
        self.ft = np.dtype(ft).type
        # Extended precision (np.longdouble, x87 80-bit): the SAME discrete problem as Float64 -- every parameter (spacings,
        # coefficients, constants) is the Float64 value, only the arithmetic is carried with 64 significand bits.  Used by
        # tests/test_oracle_extended.py to measure how far the Float64 oracle and the GPU are from the exactly-rounded answer.
        self.extended = self.ft == np.longdouble
        if self.extended:
            self.ft = np.float64
        self.topo = tuple(TOPO[t] for t in topology)
        self.N = tuple(int(n) for n in size)
        self.H = tuple(0 if self.topo[d] == FLAT else int(halo[d]) for d in range(3))
        for d in range(3):
            if self.topo[d] == FLAT:
                assert self.N[d] == 1, "Flat directions have size 1"
        self.L, self.df, self.dc, self.dfo, self.dco, self.faces, self.centers = [], [], [], [], [], [], []
        self.regular = []
        for d in range(3):
            self._generate(d, extent[d])
        if self.extended:
            self.ft = np.longdouble
            for name in ("df", "dc", "faces", "centers"):
                setattr(self, name, [a.astype(np.longdouble) for a in getattr(self, name)])
            self.L = [np.longdouble(x) for x in self.L]

    def _generate(self, d, ext):
        ft, N, H, topo = self.ft, self.N[d], self.H[d], self.topo[d]
        if topo == FLAT:
            self.L.append(ft(1)); self.df.append(np.ones(1, ft)); self.dc.append(np.ones(1, ft))
            self.dfo.append(0); self.dco.append(0); self.faces.append(np.zeros(1, ft)); self.centers.append(np.zeros(1, ft))
            self.regular.append(True)
            return
        TC = N + 2 * H
        TF = N + 2 * H + (1 if topo == BOUNDED else 0)
        if isinstance(ext, tuple) and len(ext) == 2 and np.isscalar(ext[0]):
            # regular: grid_generation.jl:117-156 (BigFloat there; longdouble here, then rounded to FT)
            c1, c2 = np.longdouble(ext[0]), np.longdouble(ext[1])
            L = c2 - c1
            D = L / N
            Fm = c1 - H * D
            Fp = Fm + (L + (2 * H - 1) * D)  # total_extent, grid_utils.jl:128
            if topo == BOUNDED:
                Fp = Fm + (L + 2 * H * D)
            F = np.linspace(ft(Fm), ft(Fp), TF).astype(ft)
            Cm = Fm + D / 2
            Cp = Cm + L + D * (2 * H - 1)
            Cc = np.linspace(ft(Cm), ft(Cp), TC).astype(ft)
            self.L.append(ft(L))
            self.df.append(np.full(TC + 1, ft(D), ft)); self.dc.append(np.full(TF, ft(D), ft))
            self.regular.append(True)
        else:
            # stretched: grid_generation.jl:34-95
            Fi = np.asarray(ext, dtype=ft)
            assert Fi.shape == (N + 1,)
            L = Fi[N] - Fi[0]
            if topo == BOUNDED:
                dlo = np.array([Fi[1] - Fi[0]] * H, ft)
                dhi = np.array([Fi[-1] - Fi[-2]] * H, ft)[::-1]
            else:
                dlo = np.array([Fi[N - H + i] - Fi[N - H + i - 1] for i in range(1, H + 1)], ft)
                dhi = np.array([Fi[i] - Fi[i - 1] for i in range(1, H + 1)], ft)[::-1]
            Flo = np.array([Fi[0] - np.sum(dlo[i:H]) for i in range(H)], ft)
            Fhi = np.array([Fi[N] + np.sum(dhi[i:H]) for i in range(H)], ft)[::-1]
            F = np.concatenate([Flo, Fi, Fhi]).astype(ft)
            Cc = np.array([(F[i + 1] + F[i]) / 2 for i in range(TC)], ft)
            dF = np.array([Cc[i] - Cc[i - 1] for i in range(1, TC)], ft)
            tF = F[:TF]
            dC = np.array([tF[i + 1] - tF[i] for i in range(TF - 1)], ft)
            dF = np.concatenate([[dF[0]], dF, [dF[-1]]])
            for i in range(len(dF) - 1, 0, -1):
                dF[i] = dF[i - 1]
            F = tF
            self.L.append(ft(L))
            self.df.append(dF.astype(ft)); self.dc.append(dC.astype(ft))
            self.regular.append(False)
        # OffsetArray offsets: Δᶠ offset -H-1 (logical i -> index i+H), Δᶜ offset -H (i -> i+H-1)
        self.dfo.append(H); self.dco.append(H - 1)
        self.faces.append(F); self.centers.append(Cc)

    # spacing lookups with logical (1-based) indices, vectorised
    def dF(self, d, idx):
        return self.df[d][np.asarray(idx) + self.dfo[d]] if self.topo[d] != FLAT else np.ones_like(np.asarray(idx), dtype=self.ft)

    def dC(self, d, idx):
        return self.dc[d][np.asarray(idx) + self.dco[d]] if self.topo[d] != FLAT else np.ones_like(np.asarray(idx), dtype=self.ft)

    def nodes(self, d, loc):
        """interior node coordinates for location 'c' or 'f' along dimension d"""
        n = self.field_size(loc, d)
        H = self.H[d]
        arr = self.centers[d] if loc == "c" else self.faces[d]
        return arr[H:H + n] if self.topo[d] != FLAT else np.zeros(1, self.ft)

    def field_size(self, loc, d):
        return self.N[d] + (1 if (loc == "f" and self.topo[d] == BOUNDED) else 0)


class Field:
    """Field{LX,LY,LZ} data (src/Fields/field.jl:21-36; new_data.jl:11-74)."""

    def __init__(self, grid, loc, bcs=None, name=""):
        self.grid, self.loc, self.name = grid, loc, name
        self.n = tuple(grid.field_size(loc[d], d) for d in range(3))  # interior size
        self.Pshape = tuple(self.n[d] + 2 * grid.H[d] for d in range(3))
        self.data = np.zeros(self.Pshape[::-1], dtype=grid.ft)
        self.bcs = bcs if bcs is not None else default_bcs(grid, loc)

    @property
    def interior(self):
        H = self.grid.H
        return self.data[H[2]:H[2] + self.n[2], H[1]:H[1] + self.n[1], H[0]:H[0] + self.n[0]]

    def view(self, i0, i1, j0, j1, k0, k1):
        """logical (1-based, inclusive) index window -> numpy view"""
        H = self.grid.H
        return self.data[k0 - 1 + H[2]:k1 + H[2], j0 - 1 + H[1]:j1 + H[1], i0 - 1 + H[0]:i1 + H[0]]

    def ofield(self):
        _, P, OField, _ = structs(self.grid.ft)
        f = OField()
        f.p = self.data.ctypes.data_as(P)
        for d in range(3):
            f.P[d] = self.Pshape[d]
            f.o[d] = self.grid.H[d]
        return f


# boundary conditions: ('periodic',), ('flux', value), ('value', v), ('gradient', g), ('impenetrable',), None
def default_bcs(grid, loc, auxiliary=False):
    """field_boundary_conditions.jl:15-36 (prognostic) / :38-52 (auxiliary: Face+Bounded -> nothing)"""
    out = {}
    for d, (lo, hi) in enumerate((("west", "east"), ("south", "north"), ("bottom", "top"))):
        t = grid.topo[d]
        if t == PERIODIC:
            bc = ("periodic",)
        elif t == FLAT:
            bc = None
        elif loc[d] == "c":
            bc = ("flux", None)  # NoFlux
        else:
            bc = None if auxiliary else ("impenetrable",)
        out[lo] = bc
        out[hi] = bc
    return out


SIDES = (("west", "east", 0), ("south", "north", 1), ("bottom", "top", 2))


def fill_halo_regions(f, fill_normal_flow_bcs=True):
    """fill_halo_regions! for one field: non-periodic sides first, periodic last
    (boundary_condition_ordering.jl:17-46,116-142).  Bit-exact restatement of the reference
    kernels fill_halo_regions_periodic.jl:5-27, _flux.jl:9-27, _value_gradient.jl:7-119,
    _normal_flow.jl:2-27, with the reference's launch extents (fill_halo_regions.jl:133-152)."""
    g = f.grid
    order = [s for s in SIDES if f.bcs[s[0]] is not None and f.bcs[s[0]][0] != "periodic"]
    # insertion sort with a non-strict `lt` reverses equal-class entries: bottom/top, south/north, west/east
    order = order[::-1]
    per = [s for s in SIDES if f.bcs[s[0]] is not None and f.bcs[s[0]][0] == "periodic"][::-1]
    for lo, hi, d in order + per:
        _fill_side_pair(f, d, f.bcs[lo], f.bcs[hi], fill_normal_flow_bcs)


def _axis_slices(f, d):
    """numpy axis for dimension d"""
    return 2 - d


def _fill_side_pair(f, d, bc_lo, bc_hi, fill_normal_flow_bcs):
    g = f.grid
    ax = 2 - d
    H, N = g.H[d], g.N[d]
    A = f.data
    if bc_lo[0] == "periodic":
        # parent(c)[i] = parent(c)[N+i]; parent(c)[N+H+i] = parent(c)[H+i], i = 1:H, over the full parent extent
        sl = [slice(None)] * 3
        src = list(sl); dst = list(sl)
        dst[ax] = slice(0, H); src[ax] = slice(N, N + H)
        A[tuple(dst)] = A[tuple(src)]
        dst[ax] = slice(N + H, N + 2 * H); src[ax] = slice(H, 2 * H)
        A[tuple(dst)] = A[tuple(src)]
        return
    # non-periodic kernels run over size(grid, loc) of the two tangential dims (interior incl. Face+1)
    t = [dd for dd in range(3) if dd != d]
    win = [slice(None)] * 3
    for dd in t:
        win[2 - dd] = slice(g.H[dd], g.H[dd] + f.n[dd])

    def plane(logical):
        w = list(win)
        w[ax] = logical - 1 + H
        return tuple(w)

    for side, bc in (("lo", bc_lo), ("hi", bc_hi)):
        kind = bc[0]
        if kind == "flux":  # c[0] = c[1] ; c[N+1] = c[N]
            if side == "lo":
                A[plane(0)] = A[plane(1)]
            else:
                A[plane(N + 1)] = A[plane(N)]
        elif kind == "impenetrable":
            if fill_normal_flow_bcs:
                A[plane(1 if side == "lo" else N + 1)] = 0
        elif kind in ("value", "gradient"):
            ft = g.ft
            # a number, or an array over the boundary plane: getbc(condition::AbstractArray, i, j, ...) = condition[i, j]
            val = np.asarray(bc[1], dtype=ft) if isinstance(bc[1], np.ndarray) else ft(bc[1])
            iB = 1 if side == "lo" else N + 1
            iI = 1 if side == "lo" else N
            iH = 0 if side == "lo" else N + 1
            # Δ at flip(loc) in the normal direction, at the boundary index
            D = g.dF(d, iB) if f.loc[d] == "c" else g.dC(d, iB)
            D = ft(D)
            cI = A[plane(iI)]
            if kind == "gradient":
                grad = val
            else:
                grad = (cI - val) / (D / 2) if side == "lo" else (val - cI) / (D / 2)
            A[plane(iH)] = cI + grad * (-D if side == "lo" else D)
        else:
            raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# Poisson solvers
# ---------------------------------------------------------------------------------------------
def _complex_of(ft):
    return np.complex128 if ft == np.float64 else np.clongdouble if ft == np.longdouble else np.complex64


def poisson_eigenvalues(grid, d):
    """poisson_eigenvalues.jl:8-32"""
    N, L, t = grid.N[d], np.float64(grid.L[d]), grid.topo[d]
    i = np.arange(N, dtype=np.float64)
    if t == PERIODIC:
        lam = (2 * np.sin(i * np.pi / N) / (L / N)) ** 2
    elif t == BOUNDED:
        lam = (2 * np.sin(i * np.pi / (2 * N)) / (L / N)) ** 2
    else:
        lam = np.zeros(N)
    return lam.astype(grid.ft)


class FFTPoissonSolver:
    """FFTBasedPoissonSolver (fft_based_poisson_solver.jl:51-73,94-124); CPU transforms: FFTW fft / REDFT10
    forward, ifft / REDFT01 * 1/(2N) backward (plan_transforms.jl:16-34, discrete_transforms.jl:26-34).
    scipy: dct(type=2) == REDFT10, idct(type=2, norm=None) == REDFT01/(2N)."""

    def __init__(self, grid):
        self.grid = grid
        self.lam = [poisson_eigenvalues(grid, d) for d in range(3)]
        self.cft = _complex_of(grid.ft)

    def solve(self, rhs):
        g = self.grid
        b = rhs.astype(self.cft)
        for d in range(3):  # Bounded first
            if g.topo[d] == BOUNDED:
                ax = 2 - d
                b = (sfft.dct(b.real, type=2, axis=ax) + 1j * sfft.dct(b.imag, type=2, axis=ax)).astype(self.cft)
        for d in range(3):
            if g.topo[d] == PERIODIC:
                b = sfft.fft(b, axis=2 - d)
        lx, ly, lz = self.lam
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = -b / (lx[None, None, :] + ly[None, :, None] + lz[:, None, None])
        phi[0, 0, 0] = 0
        for d in range(3):  # Periodic first
            if g.topo[d] == PERIODIC:
                phi = sfft.ifft(phi, axis=2 - d)
        for d in range(3):
            if g.topo[d] == BOUNDED:
                ax = 2 - d
                phi = (sfft.idct(phi.real, type=2, axis=ax) + 1j * sfft.idct(phi.imag, type=2, axis=ax)).astype(self.cft)
        return phi.real.astype(g.ft)


class FourierTridiagonalPoissonSolver:
    """FourierTridiagonalPoissonSolver, z tridiagonal (fourier_tridiagonal_poisson_solver.jl:87-149,199-260)
    + BatchedTridiagonalSolver Thomas sweep (batched_tridiagonal_solver.jl:211-243)."""

    def __init__(self, grid):
        g = self.grid = grid
        assert g.topo[2] == BOUNDED
        ft = g.ft
        self.cft = _complex_of(ft)
        Nx, Ny, Nz = g.N
        lx, ly = poisson_eigenvalues(g, 0), poisson_eigenvalues(g, 1)
        lam = (lx[None, :] + ly[:, None]).astype(ft)  # (Ny, Nx)
        k = np.arange(1, Nz + 1)
        dzf = lambda kk: g.dF(2, kk).astype(ft)
        dzc = lambda kk: g.dC(2, kk).astype(ft)
        D = np.zeros((Nz, Ny, Nx), ft)
        D[0] = ft(-1) / dzf(2) - dzc(1) * lam
        D[Nz - 1] = ft(-1) / dzf(Nz) - dzc(Nz) * lam
        for kk in range(2, Nz):
            D[kk - 1] = -(ft(1) / dzf(kk + 1) + ft(1) / dzf(kk)) - dzc(kk) * lam
        self.D = D
        self.lower = (ft(1) / dzf(np.arange(1, Nz) + 1)).astype(ft)  # lower[q] = 1/Δzᶠ(q+1), q = 1..Nz-1
        self.storage = np.zeros((Nz, Ny, Nx), self.cft)

    def solve(self, rhs):
        """rhs already multiplied by Δzᶜ (solve_for_pressure.jl:36-42)"""
        g = self.grid
        ft = g.ft
        Nx, Ny, Nz = g.N
        b = rhs.astype(self.cft)
        for d in (0, 1):
            if g.topo[d] == BOUNDED:
                ax = 2 - d
                b = (sfft.dct(b.real, type=2, axis=ax) + 1j * sfft.dct(b.imag, type=2, axis=ax)).astype(self.cft)
        for d in (0, 1):
            if g.topo[d] == PERIODIC:
                b = sfft.fft(b, axis=2 - d)
        phi, a, cdiag, D = self.storage, self.lower, self.lower, self.D
        t = np.zeros((Nz, Ny, Nx), ft)
        beta = D[0].copy()
        phi[0] = b[0] / beta
        eps10 = 10 * np.finfo(ft).eps
        for k in range(1, Nz):
            t[k] = cdiag[k - 1] / beta
            beta = D[k] - a[k - 1] * t[k]
            ok = np.abs(beta) > eps10
            with np.errstate(divide="ignore", invalid="ignore"):
                star = (b[k] - a[k - 1] * phi[k - 1]) / beta
            phi[k] = np.where(ok, star, phi[k])
        for k in range(Nz - 2, -1, -1):
            phi[k] -= t[k + 1] * phi[k + 1]
        for d in (0, 1):
            if g.topo[d] == PERIODIC:
                phi = sfft.ifft(phi, axis=2 - d)
        for d in (0, 1):
            if g.topo[d] == BOUNDED:
                ax = 2 - d
                phi = (sfft.idct(phi.real, type=2, axis=ax) + 1j * sfft.idct(phi.imag, type=2, axis=ax)).astype(self.cft)
        phi = (phi - np.mean(phi)).astype(self.cft)
        self.storage = phi
        return phi.real.astype(ft)


# ---------------------------------------------------------------------------------------------
# Model
# ---------------------------------------------------------------------------------------------
class ScalarDiffusivity:
    """ScalarDiffusivity(time_discretization, ν, κ): vertically_implicit = VerticallyImplicitTimeDiscretization()
    (scalar_diffusivity.jl:113-137)"""

    def __init__(self, nu=0.0, kappa=0.0, vertically_implicit=False):
        self.nu, self.kappa, self.vertically_implicit = nu, kappa, bool(vertically_implicit)


class Smagorinsky:
    """Smagorinsky([time_discretization]; coefficient, Pr) (smagorinsky.jl:76-84); vertically_implicit =
    VerticallyImplicitTimeDiscretization(): the implicit step takes nu_e interpolated to the nodes of its coefficients"""

    def __init__(self, coefficient=0.16, Pr=1.0, lilly=False, Cb=1.0, vertically_implicit=False, dynamic=None):
        self.cs, self.Pr, self.lilly, self.cb = coefficient, Pr, lilly, Cb
        self.vertically_implicit = bool(vertically_implicit)
        # DynamicCoefficient(averaging = dims; minimum_numerator) (dynamic_coefficient.jl:208-212): {"averaging": (1, 2), ...}
        self.dynamic = None if dynamic is None else {"averaging": tuple(dynamic.get("averaging", (1, 2))),
                                                     "minimum_numerator": dynamic.get("minimum_numerator", 1e-32)}


def DynamicSmagorinsky(averaging=(1, 2), Pr=1.0, minimum_numerator=1e-32, vertically_implicit=False):
    """DynamicSmagorinsky(; averaging, Pr, minimum_numerator) with a directional average (dynamic_coefficient.jl:107-118)"""
    averaging = (averaging,) if isinstance(averaging, int) else tuple(averaging)
    return Smagorinsky(coefficient=0.0, Pr=Pr, vertically_implicit=vertically_implicit,
                       dynamic={"averaging": averaging, "minimum_numerator": minimum_numerator})


def SmagorinskyLilly(C=0.16, Cb=1.0, Pr=1.0, vertically_implicit=False):
    return Smagorinsky(coefficient=C, Pr=Pr, lilly=True, Cb=Cb, vertically_implicit=vertically_implicit)


class AnisotropicMinimumDissipation:
    """AnisotropicMinimumDissipation([time_discretization]; C, Cν, Cκ, Cb) (anisotropic_minimum_dissipation.jl:124-139)"""

    def __init__(self, C=1.0 / 3.0, Cnu=None, Ckappa=None, Cb=None, vertically_implicit=False):
        self.Cnu = C if Cnu is None else Cnu
        self.Ckappa = C if Ckappa is None else Ckappa
        self.Cb = Cb
        self.vertically_implicit = bool(vertically_implicit)


class Model:
    """NonhydrostaticModel restatement (nonhydrostatic_model.jl:124-313)."""

    def __init__(self, grid, advection=("centered", 2), closure=None, buoyancy=None, coriolis_f=None,
                 tracers=(), timestepper="rk3", boundary_conditions=None, chi=0.1):
        self.grid = g = grid
        ft = g.ft
        self.advection = advection
        self.closures = [] if closure is None else (list(closure) if isinstance(closure, (tuple, list)) else [closure])
        self.buoyancy = buoyancy  # None | ('tracer',) | ('seawater', g, alpha, beta)
        self.coriolis_f = coriolis_f
        self.tracer_names = tuple(tracers)
        self.timestepper = timestepper
        self.chi = ft(chi)
        bcs = boundary_conditions or {}

        def mk(name, loc, aux=False):
            b = default_bcs(g, loc, auxiliary=aux)
            b.update(bcs.get(name, {}))
            return Field(g, loc, b, name)

        self.u, self.v, self.w = mk("u", "fcc"), mk("v", "cfc"), mk("w", "ccf")
        self.tracers = [mk(n, "ccc") for n in self.tracer_names]
        self.pNHS = mk("pNHS", "ccc")
        self.pHY = mk("pHY", "ccc") if (buoyancy is not None and g.topo[2] != PERIODIC) else None
        names = ("u", "v", "w") + self.tracer_names
        locs = ("fcc", "cfc", "ccf") + ("ccc",) * len(self.tracers)
        self.Gn = [Field(g, l, name="Gn_" + n) for n, l in zip(names, locs)]
        self.Gm = [Field(g, l, name="Gm_" + n) for n, l in zip(names, locs)]
        self.dynamic_fields = {}   # DynamicSmagorinsky: closure index -> {Sigma, Sigmabar, LM, MM, JLM, JMM} (oracle/dynsmag.py)
        self.nue = [mk("nue%d" % m, "ccc", aux=True) if not isinstance(c, ScalarDiffusivity) else None
                    for m, c in enumerate(self.closures)]
        self.kappae = [[mk("kappae%d_%s" % (m, n), "ccc", aux=True) for n in self.tracer_names]
                       if isinstance(c, AnisotropicMinimumDissipation) else None for m, c in enumerate(self.closures)]
        if all(g.regular[d] for d in range(3)):
            self.solver = FFTPoissonSolver(g)
            self.tridiagonal = False
        else:
            assert g.regular[0] and g.regular[1], "only z may be stretched"
            self.solver = FourierTridiagonalPoissonSolver(g)
            self.tridiagonal = True
        self.time = 0.0
        self.iteration = 0
        self.last_dt = np.inf
        tft = np.float64 if ft == np.longdouble else ft   # extended precision keeps the Float64 coefficient VALUES
        self._tables = tuple(t.astype(ft) for t in (coef.weno_coeff_table(tft), coef.smoothness_table(tft), coef.cstar_table(tft),
                                                    coef.centered_coeff_table(tft)))
        self.update_state()

    # -- helpers ------------------------------------------------------------------------------
    @property
    def prognostic(self):
        return [self.u, self.v, self.w] + self.tracers

    def params(self):
        g = self.grid
        ft = g.ft
        cft, P, OField, OParams = structs(ft)
        p = OParams()
        self._keep = []
        for d in range(3):
            p.N[d], p.H[d], p.topo[d] = g.N[d], g.H[d], g.topo[d]
            p.df[d] = g.df[d].ctypes.data_as(P); p.dc[d] = g.dc[d].ctypes.data_as(P)
            p.dfo[d], p.dco[d] = g.dfo[d], g.dco[d]
        adv = self.advection
        if adv is None:
            p.adv_kind, p.adv_buffer = 0, 0
        elif adv[0] == "weno":
            p.adv_kind, p.adv_buffer = 2, (adv[1] + 1) // 2
        else:
            p.adv_kind, p.adv_buffer = 1, adv[1] // 2
        wc, wb, cs, cc = self._tables
        p.weno_coeff = wc.ctypes.data_as(P); p.weno_beta = wb.ctypes.data_as(P)
        p.weno_cstar = cs.ctypes.data_as(P); p.cen_coeff = cc.ctypes.data_as(P)
        p.weno_eps = ft(coef.weno_eps(np.float64 if ft == np.longdouble else ft))
        p.nclosures = len(self.closures)
        nt = len(self.tracers)
        for m, c in enumerate(self.closures):
            if isinstance(c, ScalarDiffusivity):
                p.closure_kind[m] = 1
                p.closure_vi[m] = int(c.vertically_implicit)
                p.nu[m] = ft(c.nu)
                for t in range(nt):
                    p.kappa[m][t] = ft(_per_tracer(c.kappa, self.tracer_names, t))
            elif isinstance(c, Smagorinsky):
                p.closure_kind[m] = 2
                p.closure_vi[m] = int(c.vertically_implicit)
                p.cs[m], p.cb[m], p.lilly[m] = ft(c.cs), ft(c.cb), int(c.lilly)
                for t in range(nt):
                    p.Pr[m][t] = ft(_per_tracer(c.Pr, self.tracer_names, t))
                p.nue[m] = self.nue[m].ofield()
            else:
                p.closure_kind[m] = 3
                p.closure_vi[m] = int(c.vertically_implicit)
                p.Cnu[m] = ft(c.Cnu)
                p.amd_has_cb[m] = 0 if c.Cb is None else 1
                p.cb[m] = ft(0 if c.Cb is None else c.Cb)
                for t in range(nt):
                    p.Ckappa[m][t] = ft(_per_tracer(c.Ckappa, self.tracer_names, t))
                    p.kappae[m][t] = self.kappae[m][t].ofield()
                p.nue[m] = self.nue[m].ofield()
        b = self.buoyancy
        if b is None:
            p.buoy_kind = 0
        elif b[0] == "tracer":
            p.buoy_kind, p.ib = 1, self.tracer_names.index("b")
        else:
            p.buoy_kind, p.iT, p.iS = 2, self.tracer_names.index("T"), self.tracer_names.index("S")
            p.g, p.alpha, p.beta = ft(b[1]), ft(b[2]), ft(b[3])
        p.has_coriolis = 0 if self.coriolis_f is None else 1
        p.fcor = ft(self.coriolis_f or 0)
        p.ntracers = nt
        p.u, p.v, p.w = self.u.ofield(), self.v.ofield(), self.w.ofield()
        for t in range(nt):
            p.c[t] = self.tracers[t].ofield()
            p.Gc[t] = self.Gn[3 + t].ofield()
        p.has_pHY = 0 if self.pHY is None else 1
        if self.pHY is not None:
            p.pHY = self.pHY.ofield()
        p.Gu, p.Gv, p.Gw = self.Gn[0].ofield(), self.Gn[1].ofield(), self.Gn[2].ofield()
        return p

    def _fn(self, name):
        return getattr(lib(), name + ("_f64" if self.grid.ft == np.float64 else "_f80" if self.grid.ft == np.longdouble else "_f32"))

    # -- update_state! (update_nonhydrostatic_model_state.jl:22-83) -----------------------------
    def update_state(self):
        for f in self.prognostic:
            fill_halo_regions(f, fill_normal_flow_bcs=False)
        self.compute_closure_fields()
        self.update_hydrostatic_pressure()
        for m in range(len(self.closures)):
            if self.nue[m] is not None:
                fill_halo_regions(self.nue[m])
            if self.kappae[m] is not None:
                for k in self.kappae[m]:
                    fill_halo_regions(k)
        if self.pHY is not None:
            fill_halo_regions(self.pHY)
        self.compute_tendencies()

    def compute_closure_fields(self):
        p = None
        for m, c in enumerate(self.closures):
            if isinstance(c, Smagorinsky) and c.dynamic is not None:
                from .dynsmag import compute_dynamic_smagorinsky
                compute_dynamic_smagorinsky(self, m)
            elif isinstance(c, Smagorinsky):
                p = p or self.params()
                self._fn("orc_smagorinsky_viscosity")(C.byref(p), m)
            elif isinstance(c, AnisotropicMinimumDissipation):
                p = p or self.params()
                self._fn("orc_amd_viscosity")(C.byref(p), m)
                for t in range(len(self.tracers)):
                    self._fn("orc_amd_diffusivity")(C.byref(p), m, t)

    def buoyancy_perturbation(self):
        b = self.buoyancy
        ft = self.grid.ft
        if b[0] == "tracer":
            return self.tracers[self.tracer_names.index("b")].data
        T = self.tracers[self.tracer_names.index("T")].data
        S = self.tracers[self.tracer_names.index("S")].data
        return (ft(b[1]) * (ft(b[2]) * T - ft(b[3]) * S)).astype(ft)

    def update_hydrostatic_pressure(self):
        """update_hydrostatic_pressure.jl:11-39 over i,j in (-H+2 : N+H-1) (interleave…jl:85-94)"""
        if self.pHY is None or self.grid.topo[2] == FLAT:
            return
        g = self.grid
        ft = g.ft
        Nz, Hz = g.N[2], g.H[2]
        b = self.buoyancy_perturbation()
        A = self.pHY.data
        sl = []
        for d in (1, 0):  # y then x (numpy axes 1, 2)
            if g.topo[d] == FLAT:
                sl.append(slice(None))
            else:
                sl.append(slice(1, g.N[d] + 2 * g.H[d] - 1))  # logical -H+2 .. N+H-1
        sl = tuple(sl)

        def bf(k):  # ℑzᵃᵃᶠ(b) at face k : 0.5*(b[k-1] + b[k])
            return ft(0.5) * (b[(k - 2 + Hz,) + sl] + b[(k - 1 + Hz,) + sl])

        pk = -bf(Nz + 1) * ft(g.dF(2, Nz + 1))
        A[(Nz - 1 + Hz,) + sl] = pk
        for k in range(Nz - 1, 0, -1):
            pk = pk - bf(k + 1) * ft(g.dF(2, k + 1))
            A[(k - 1 + Hz,) + sl] = pk

    def compute_tendencies(self):
        p = self.params()
        self._fn("orc_compute_tendencies")(C.byref(p))

    # -- flux boundary conditions (compute_flux_bcs.jl:113-162) ---------------------------------
    def compute_flux_bc_tendencies(self):
        g = self.grid
        ft = g.ft
        for f, G in zip(self.prognostic, self.Gn):
            for lo, hi, d in SIDES:
                for side, name in (("lo", lo), ("hi", hi)):
                    bc = f.bcs[name]
                    if bc is None or bc[0] != "flux" or bc[1] is None:
                        continue
                    if isinstance(bc[1], np.ndarray):   # array-valued flux over the field's tangential extent; kernel over size(grid)
                        t0, t1 = [dd for dd in range(3) if dd != d]
                        flux = np.expand_dims(np.asarray(bc[1], dtype=ft)[:g.N[t1], :g.N[t0]], axis=2 - d)
                    else:
                        flux = ft(bc[1])
                    self._add_flux(f, G, d, side, flux)

    def _add_flux(self, f, G, d, side, flux):
        g = self.grid
        ft = g.ft
        N = g.N
        idx = [np.arange(1, N[dd] + 1) for dd in range(3)]  # launched over :yz etc. => size(grid)
        n_int = 1 if side == "lo" else N[d]
        n_face = 1 if side == "lo" else N[d] + 1

        def sp(dd, loc, ii):
            return (g.dF(dd, ii) if loc == "f" else g.dC(dd, ii)).astype(ft)

        flip = lambda l: "f" if l == "c" else "c"
        loc = f.loc
        t = [dd for dd in range(3) if dd != d]
        # area normal to d at flipped location, at index n_face
        a0 = sp(t[0], loc[t[0]], idx[t[0]])
        a1 = sp(t[1], loc[t[1]], idx[t[1]])
        # volume at (loc) at index n_int: V = (Δx*Δy)*Δz
        dn = sp(d, loc[d], np.array([n_int]))[0]
        shape = [1, 1, 1]
        A0 = a0.reshape([-1 if (2 - t[0]) == ax else 1 for ax in range(3)])
        A1 = a1.reshape([-1 if (2 - t[1]) == ax else 1 for ax in range(3)])
        if d == 0:      # Ax = Δy*Δz ; V = (Δx*Δy)*Δz
            area = A0 * A1
            vol = (dn * A0) * A1
        elif d == 1:    # Ay = Δx*Δz
            area = A0 * A1
            vol = (A0 * dn) * A1
        else:           # Az = Δx*Δy
            area = A0 * A1
            vol = (A0 * A1) * dn
        win = [None, None, None]
        H = g.H
        for dd in t:
            win[2 - dd] = slice(H[dd], H[dd] + N[dd])
        win[2 - d] = slice(n_int - 1 + H[d], n_int + H[d])
        sl = tuple(win)
        term = (flux * area / vol).astype(ft)
        if side == "lo":
            G.data[sl] += term
        else:
            G.data[sl] -= term

    # -- time stepping ----------------------------------------------------------------------------
    def _periphery_window(self, f, n):
        """:xyz minus boundary-normal faces for u, v, w (kernel_launching.jl:160-173,259-268)"""
        g = self.grid
        lo = [1, 1, 1]
        hi = list(g.N)
        if n < 3 and g.topo[n] == BOUNDED:
            lo[n] = 2
        return f.view(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def _window_like(self, f, n, other):
        g = self.grid
        lo = [1, 1, 1]
        hi = list(g.N)
        if n < 3 and g.topo[n] == BOUNDED:
            lo[n] = 2
        return other.view(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def rk3_substep(self, dt, gamma, zeta):
        ft = self.grid.ft
        self.compute_flux_bc_tendencies()
        self._dt_user = float(dt)
        dt = ft(dt)
        for n, f in enumerate(self.prognostic):
            U = self._periphery_window(f, n)
            Gn = self._window_like(f, n, self.Gn[n])
            Gm = self._window_like(f, n, self.Gm[n])
            if zeta is None:
                U += dt * ft(gamma) * Gn
            else:
                U += dt * (ft(gamma) * Gn + ft(zeta) * Gm)
        # Δτ = convert(FT, stage_Δt(Δt, γ, ζ)) with stage_Δt = Δt * (γ + ζ) (runge_kutta_3.jl:186-187)
        gz = ft(gamma) if zeta is None else ft(ft(gamma) + ft(zeta))
        dtau = ft(self._dt_user * float(gz))
        # the reference interleaves (explicit update, implicit_step!) field by field; both are column-local per field
        self.implicit_step(dtau)
        self.pressure_correct(dtau)

    def implicit_step(self, dtau):
        """implicit_step! of every prognostic field (nonhydrostatic_rk3_substep.jl:47-56, nonhydrostatic_ab2_step.jl:41-50)"""
        if not any(getattr(c, "vertically_implicit", False) for c in self.closures):
            return
        if self.grid.topo[2] != BOUNDED:
            raise ValueError("VerticallyImplicitTimeDiscretization can only be specified on grids that are Bounded in the z-direction.")
        ft = self.grid.ft
        p = self.params()
        scratch = np.zeros(self.grid.N[2] + 2, dtype=ft)
        fn = self._fn("orc_implicit_step")
        fn.restype = None
        cft = structs(ft)[0]
        for n, f in enumerate(self.prognostic):
            of = f.ofield()
            fn(C.byref(p), C.c_int(n), C.byref(of), cft(float(dtau)), scratch.ctypes.data_as(C.c_void_p))

    def ab2_step(self, dt, chi):
        ft = self.grid.ft
        self.compute_flux_bc_tendencies()
        dt = ft(dt)
        alpha = ft(1.5) + chi
        beta = ft(0.5) + chi
        not_euler = ft(1) if chi != ft(-0.5) else ft(0)
        for n, f in enumerate(self.prognostic):
            U = self._periphery_window(f, n)
            Gn = self._window_like(f, n, self.Gn[n])
            Gm = self._window_like(f, n, self.Gm[n])
            # `β * G⁻ * not_euler`: a false Bool is a strong zero in Julia (kills leftover NaNs in G⁻)
            Gu = alpha * Gn - beta * Gm if not_euler else alpha * Gn - ft(0)
            U += dt * Gu
        self.implicit_step(dt)
        self.pressure_correct(dt)

    def divergence(self):
        """divᶜᶜᶜ (divergence_operators.jl:16-19) on the interior"""
        g = self.grid
        ft = g.ft
        Nx, Ny, Nz = g.N
        i = np.arange(1, Nx + 1); j = np.arange(1, Ny + 1); k = np.arange(1, Nz + 1)
        dxc = g.dC(0, i).astype(ft)[None, None, :]; dyc = g.dC(1, j).astype(ft)[None, :, None]; dzc = g.dC(2, k).astype(ft)[:, None, None]
        Ax = dyc * dzc; Ay = dxc * dzc; Az = dxc * dyc
        u, v, w = self.u, self.v, self.w
        zero = ft(0)
        dx = (Ax * u.view(2, Nx + 1, 1, Ny, 1, Nz) - Ax * u.view(1, Nx, 1, Ny, 1, Nz)) if g.topo[0] != FLAT else zero
        dy = (Ay * v.view(1, Nx, 2, Ny + 1, 1, Nz) - Ay * v.view(1, Nx, 1, Ny, 1, Nz)) if g.topo[1] != FLAT else zero
        dz = (Az * w.view(1, Nx, 1, Ny, 2, Nz + 1) - Az * w.view(1, Nx, 1, Ny, 1, Nz)) if g.topo[2] != FLAT else zero
        Vinv = ft(1) / ((dxc * dyc) * dzc)
        return (Vinv * (dx + dy + dz)).astype(ft), dzc

    def pressure_correct(self, dtau):
        """compute_pressure_correction! + make_pressure_correction! (pressure_correction.jl:6-106)"""
        g = self.grid
        ft = g.ft
        for f in (self.u, self.v, self.w):
            fill_halo_regions(f)
        div, dzc = self.divergence()
        if self.tridiagonal:
            rhs = (dzc * div).astype(ft)
        else:
            rhs = div
        p = self.solver.solve(rhs)
        self.pNHS.interior[...] = p
        fill_halo_regions(self.pNHS)
        Nx, Ny, Nz = g.N
        P = self.pNHS
        i = np.arange(1, Nx + 1); j = np.arange(1, Ny + 1); k = np.arange(1, Nz + 1)
        if g.topo[0] != FLAT:
            rdx = (ft(1) / g.dF(0, i).astype(ft))[None, None, :]
            self.u.view(1, Nx, 1, Ny, 1, Nz)[...] -= (P.view(1, Nx, 1, Ny, 1, Nz) - P.view(0, Nx - 1, 1, Ny, 1, Nz)) * rdx
        if g.topo[1] != FLAT:
            rdy = (ft(1) / g.dF(1, j).astype(ft))[None, :, None]
            self.v.view(1, Nx, 1, Ny, 1, Nz)[...] -= (P.view(1, Nx, 1, Ny, 1, Nz) - P.view(1, Nx, 0, Ny - 1, 1, Nz)) * rdy
        if g.topo[2] != FLAT:
            rdz = (ft(1) / g.dF(2, k).astype(ft))[:, None, None]
            self.w.view(1, Nx, 1, Ny, 1, Nz)[...] -= (P.view(1, Nx, 1, Ny, 1, Nz) - P.view(1, Nx, 1, Ny, 0, Nz - 1)) * rdz
        P.data /= max(np.finfo(ft).eps, ft(dtau))

    def cache_previous_tendencies(self):
        for Gn, Gm in zip(self.Gn, self.Gm):
            N = self.grid.N  # launched over :xyz == size(grid) (cache_nonhydrostatic_tendencies.jl:21-31)
            Gm.view(1, N[0], 1, N[1], 1, N[2])[...] = Gn.view(1, N[0], 1, N[1], 1, N[2])

    def time_step(self, dt, euler=False):
        ft = self.grid.ft
        if self.timestepper == "rk3":
            g1, g2, g3 = ft(8.0 / 15.0), ft(5.0 / 12.0), ft(3.0 / 4.0)
            z2, z3 = ft(-17.0 / 60.0), ft(-5.0 / 12.0)
            if self.iteration == 0:
                self.update_state()
            for gam, zet in ((g1, None), (g2, z2), (g3, z3)):
                self.rk3_substep(dt, gam, zet)
                self.cache_previous_tendencies()
                self.update_state()
        else:
            euler = euler or (dt != self.last_dt)
            if self.iteration == 0:
                self.update_state()
            chi = ft(-0.5) if euler else self.chi
            self.ab2_step(dt, chi)
            self.cache_previous_tendencies()
            self.update_state()
        self.last_dt = dt
        self.time += dt
        self.iteration += 1

    def set(self, enforce_incompressibility=True, **kw):
        """set!(model; ...) (set_nonhydrostatic_model.jl:39-74)"""
        byname = {"u": self.u, "v": self.v, "w": self.w}
        byname.update({n: f for n, f in zip(self.tracer_names, self.tracers)})
        for name, val in kw.items():
            f = byname[name]
            f.interior[...] = np.asarray(val, dtype=self.grid.ft).reshape(f.interior.shape)
            fill_halo_regions(f)
        self.update_state()
        if enforce_incompressibility:
            self.pressure_correct(self.grid.ft(1))
            self.update_state()


def _per_tracer(v, names, t):
    if isinstance(v, dict):
        return v[names[t]]
    if isinstance(v, (tuple, list)):
        return v[t]
    return v
