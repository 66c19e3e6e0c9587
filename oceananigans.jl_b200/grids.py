"""Host-side mirror of `B200()` and `RectilinearGrid(arch, FT; size, x, y, z, extent, halo, topology)`.

Reference: src/Architectures.jl:21-132 (architecture), src/Grids/rectilinear_grid.jl:266-294 (constructor),
src/Grids/grid_generation.jl:34-156 (coordinate generation), src/Grids/input_validation.jl:71-93 (halo default
(3,3,3) clipped to the size).  The grid lives on the host (as it does in Julia); only the spacings cross the
C ABI (ob_grid_desc).
"""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _abi

Float64 = np.float64
Float32 = np.float32


class Periodic:
    pass


class Bounded:
    pass


class Flat:
    pass


_TOPO = {Periodic: _abi.OB_PERIODIC, Bounded: _abi.OB_BOUNDED, Flat: _abi.OB_FLAT,
         "Periodic": _abi.OB_PERIODIC, "Bounded": _abi.OB_BOUNDED, "Flat": _abi.OB_FLAT}


class B200:
    """The new architecture (`const B200 = GPU{B200Device}` in the Julia extension).  Owns a library context
    (device + stream).  `B200(device)`; with `torch.distributed` initialised, `Distributed(B200())` adds the
    slab-x communicator (see distributed.py)."""

    def __init__(self, device=0):
        self.device = int(device)
        h = C.c_void_p()
        _abi.call("ob_init", self.device, C.byref(h))
        self.ctx = h
        self.rank, self.world = 0, 1

    def synchronize(self):
        _abi.call("ob_sync", self.ctx)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                _abi.lib().ob_shutdown(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def __repr__(self):
        return "B200(device=%d)" % self.device


class CPU:
    def __init__(self, *a, **k):
        raise NotImplementedError("ocean_b200 has no CPU path: use B200() (the reference's CPU() lives in Oceananigans.jl)")


def _inflate(tup, topo, default):
    """(a, b) for a grid with one Flat direction -> 3-tuple (input_validation.jl inflate_tuple)"""
    tup = (tup,) if np.isscalar(tup) else tuple(tup)
    n_nonflat = sum(t != _abi.OB_FLAT for t in topo)
    if len(tup) == 3:
        return tuple(tup)
    if len(tup) != n_nonflat:
        raise ValueError("length(%r) must be %d for this topology" % (tup, n_nonflat))
    it = iter(tup)
    return tuple(default if t == _abi.OB_FLAT else next(it) for t in topo)


class RectilinearGrid:
    def __init__(self, architecture=None, FT=Float64, *, size, x=None, y=None, z=None, halo=None, extent=None,
                 topology=(Periodic, Periodic, Bounded)):
        if architecture is None:
            raise ValueError("RectilinearGrid needs an architecture: RectilinearGrid(B200(), size=...)")
        self.architecture = architecture
        self.FT = np.dtype(FT).type
        self.topology = tuple(topology)
        topo = self.topo = tuple(_TOPO[t] for t in topology)
        self.N = tuple(int(n) for n in _inflate(size, topo, 1))
        # Distributed(B200()): `size` and the x interval are GLOBAL; this rank owns an equal slab in x
        # (DistributedComputations/distributed_grids.jl: local size = global size / partition)
        self.world, self.rank = getattr(architecture, "world", 1), getattr(architecture, "rank", 0)
        self.N_global = self.N
        self._x0 = 0
        if self.world > 1:
            from .distributed import partition_x
            if topo[0] != _abi.OB_PERIODIC:
                raise _abi.OceanB200Error(-3, "slab-x distributed grids need a Periodic x")
            nx, self._x0 = partition_x(self.N[0], self.world, self.rank)
            self.N = (nx,) + self.N[1:]
        if halo is None:
            halo = tuple(min(3, n) for n in self.N)  # validate_halo(::Nothing)
            halo = tuple(h for h, t in zip(halo, topo) if t != _abi.OB_FLAT)
        H = _inflate(halo, topo, 0)
        self.H = tuple(0 if t == _abi.OB_FLAT else int(h) for h, t in zip(H, topo))
        for d in range(3):
            if topo[d] == _abi.OB_FLAT and self.N[d] != 1:
                raise ValueError("Flat dimensions have size 1")
            if self.H[d] > self.N[d]:
                raise ValueError("halo=%d must be <= size=%d" % (self.H[d], self.N[d]))
        if extent is not None:
            if any(c is not None for c in (x, y, z)):
                raise ValueError("Cannot specify both 'extent' and 'x, y, z' keyword arguments.")
            ext = _inflate(extent, topo, None)
            coords = [None if e is None else (0, e) for e in ext]
            # validate_rectilinear_domain: z = (-Lz, 0)
            if coords[2] is not None:
                coords[2] = (-ext[2], 0)
        else:
            coords = [x, y, z]
        self._coords = coords
        self.L, self.faces, self.centers, self.dF, self.dC, self.regular = [], [], [], [], [], []
        for d in range(3):
            self._generate(d, coords[d])
        self.Nx, self.Ny, self.Nz = self.N
        self.Hx, self.Hy, self.Hz = self.H
        self.Lx, self.Ly, self.Lz = self.L

    # grid_generation.jl:34-156 ------------------------------------------------------------------------------
    def _generate(self, d, c):
        ft, N, H, topo = self.FT, self.N[d], self.H[d], self.topo[d]
        split = d == 0 and self.world > 1
        if split:
            N = self.N_global[0]   # generate the global coordinate, then keep this rank's window (same Δx on every rank)
        if topo == _abi.OB_FLAT:
            self.L.append(ft(1)); self.faces.append(np.zeros(1, ft)); self.centers.append(np.zeros(1, ft))
            self.dF.append(ft(1)); self.dC.append(ft(1)); self.regular.append(True)
            return
        if c is None:
            raise ValueError("coordinate %s of a non-Flat dimension must be given" % "xyz"[d])
        TC = N + 2 * H
        TF = TC + (1 if topo == _abi.OB_BOUNDED else 0)
        if callable(c):
            c = [c(k) for k in range(1, N + 2)]
        if isinstance(c, tuple) and len(c) == 2:
            # regular: BigFloat arithmetic in the reference; exact rationals here, rounded once to FT
            c1, c2 = Fraction(float(c[0])), Fraction(float(c[1]))
            if not c1 < c2:
                raise ValueError("%s must be an increasing interval" % "xyz"[d])
            L = c2 - c1
            D = L / N
            Fm = c1 - H * D
            Fp = Fm + (L + (2 * H - 1) * D if topo == _abi.OB_PERIODIC else L + 2 * H * D)
            Cm = Fm + D / 2
            Cp = Cm + L + D * (2 * H - 1)
            rnd = (lambda q: ft(float(q))) if ft == np.float64 else (lambda q: np.float32(float(q)))
            F, Cc = np.linspace(rnd(Fm), rnd(Fp), TF).astype(ft), np.linspace(rnd(Cm), rnd(Cp), TC).astype(ft)
            if split:
                n = self.N[0]
                F, Cc, L = F[self._x0:self._x0 + n + 2 * H], Cc[self._x0:self._x0 + n + 2 * H], L / self.world
            self.faces.append(F); self.centers.append(Cc)
            self.L.append(rnd(L)); self.dF.append(rnd(D)); self.dC.append(rnd(D)); self.regular.append(True)
            return
        if split:
            raise _abi.OceanB200Error(-3, "a partitioned x must be regularly spaced")
        # variably spaced: explicit interior faces
        Fi = np.asarray(c, dtype=ft)
        if Fi.shape != (N + 1,):
            raise ValueError("length(%s) must be N+1 = %d" % ("xyz"[d], N + 1))
        if np.any(np.diff(Fi) <= 0):
            raise ValueError("The elements of %s must be increasing!" % "xyz"[d])
        if topo == _abi.OB_BOUNDED:
            dlo = np.full(H, Fi[1] - Fi[0], ft)
            dhi = np.full(H, Fi[N] - Fi[N - 1], ft)
        else:
            dlo = np.array([Fi[N - H + i] - Fi[N - H + i - 1] for i in range(1, H + 1)], ft)
            dhi = np.array([Fi[i] - Fi[i - 1] for i in range(1, H + 1)], ft)[::-1]
        Flo = np.array([Fi[0] - np.sum(dlo[i:H], dtype=ft) for i in range(H)], ft)
        Fhi = np.array([Fi[N] + np.sum(dhi[i:H], dtype=ft) for i in range(H)], ft)[::-1]
        F = np.concatenate([Flo, Fi, Fhi]).astype(ft)
        Cc = ((F[1:TC + 1] + F[:TC]) / ft(2)).astype(ft)
        dFv = (Cc[1:] - Cc[:-1]).astype(ft)
        tF = F[:TF]
        dCv = (tF[1:] - tF[:-1]).astype(ft)
        dFv = np.concatenate([[dFv[0]], dFv, [dFv[-1]]]).astype(ft)
        dFv[1:] = dFv[:-1].copy()
        self.L.append(ft(Fi[N] - Fi[0])); self.faces.append(tF); self.centers.append(Cc)
        # OffsetArrays: Δᶠ[i] = dFv[i + H] ; Δᶜ[i] = dCv[i + H - 1]
        self.dF.append(np.ascontiguousarray(dFv)); self.dC.append(np.ascontiguousarray(dCv)); self.regular.append(False)

    # ---------------------------------------------------------------------------------------------------------
    def with_halo(self, halo):
        """with_halo(new_halo, grid) (rectilinear_grid.jl:440-455)"""
        kw = dict(size=tuple(n for n, t in zip(self.N_global, self.topo) if t != _abi.OB_FLAT),
                  halo=tuple(h for h, t in zip(halo, self.topo) if t != _abi.OB_FLAT), topology=self.topology)
        xyz = {}
        for d, name in enumerate("xyz"):
            if self.topo[d] == _abi.OB_FLAT:
                continue
            if self.regular[d]:
                xyz[name] = self._coords[d]
            else:
                H = self.H[d]
                xyz[name] = np.array(self.faces[d][H:H + self.N[d] + 1])
        return RectilinearGrid(self.architecture, self.FT, **kw, **xyz)

    def nodes(self, d, loc):
        """interior nodes along dimension d at location 'c' / 'f' (xnodes, ynodes, znodes)"""
        if self.topo[d] == _abi.OB_FLAT:
            return np.zeros(1, self.FT)
        H = self.H[d]
        n = self.N[d] + (1 if (loc == "f" and self.topo[d] == _abi.OB_BOUNDED) else 0)
        return (self.centers[d] if loc == "c" else self.faces[d])[H:H + n]

    def desc(self):
        g = _abi.GridDesc()
        g.float_type = _abi.OB_F64 if self.FT == np.float64 else _abi.OB_F32
        for d in range(3):
            g.N[d], g.H[d], g.topology[d] = self.N[d], self.H[d], self.topo[d]
            g.L[d] = float(self.L[d])
            g.d[d] = float(self.dF[d]) if self.regular[d] else 0.0
        if not (self.regular[0] and self.regular[1]):
            raise _abi.OceanB200Error(-3, "only z may be variably spaced (FFT / Fourier-tridiagonal pressure solvers)")
        if not self.regular[2]:
            g.dzf_host = self.dF[2].ctypes.data_as(C.c_void_p)
            g.dzc_host = self.dC[2].ctypes.data_as(C.c_void_p)
            g.n_dzf, g.n_dzc = len(self.dF[2]), len(self.dC[2])
        return g

    def __repr__(self):
        return "%dx%dx%d RectilinearGrid{%s} on %r with %dx%dx%d halo" % (*self.N, self.FT.__name__, self.architecture, *self.H)
