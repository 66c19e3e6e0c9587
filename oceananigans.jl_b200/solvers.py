"""Stand-alone pressure solvers (host mirror of src/Solvers): `FFTBasedPoissonSolver(grid)`,
`FourierTridiagonalPoissonSolver(grid)`, `BatchedTridiagonalSolver`, `solve!(ϕ, solver, rhs)` -> `solve`.

Reference: fft_based_poisson_solver.jl:51-124, fourier_tridiagonal_poisson_solver.jl:87-260,
batched_tridiagonal_solver.jl:9-243.  The model owns its solver inside the library; these wrappers exist for the
solver-level parity tests (test/test_poisson_solvers.jl, test_batched_tridiagonal_solver.jl).
"""
import ctypes as C

import numpy as np

from . import _abi


class _DeviceBuffer:
    def __init__(self, arch, nbytes):
        self.arch, self.nbytes = arch, nbytes
        p = C.c_void_p()
        _abi.call("ob_malloc", arch.ctx, nbytes, C.byref(p))
        self.ptr = p

    def upload(self, a):
        a = np.ascontiguousarray(a)
        assert a.nbytes == self.nbytes
        _abi.call("ob_memcpy_h2d", self.arch.ctx, self.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes)
        _abi.call("ob_sync", self.arch.ctx)

    def download(self, shape, dtype):
        out = np.empty(shape, dtype)
        assert out.nbytes == self.nbytes
        _abi.call("ob_memcpy_d2h", self.arch.ctx, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes)
        return out

    def __del__(self):
        try:
            if self.ptr and getattr(self.arch, "ctx", None):
                _abi.lib().ob_free(self.arch.ctx, self.ptr)
                self.ptr = None
        except Exception:
            pass


class _PoissonSolver:
    def __init__(self, grid):
        self.grid = grid
        self.arch = grid.architecture
        self._desc = grid.desc()
        h = C.c_void_p()
        _abi.call("ob_solver_create", self.arch.ctx, C.byref(self._desc), C.byref(h))
        self.handle = h
        n = grid.N[0] * grid.N[1] * grid.N[2] * np.dtype(grid.FT).itemsize
        self._rhs = _DeviceBuffer(self.arch, n)
        self._phi = _DeviceBuffer(self.arch, n)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _abi.lib().ob_solver_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class FFTBasedPoissonSolver(_PoissonSolver):
    def __init__(self, grid):
        if not all(grid.regular):
            raise ValueError("FFTBasedPoissonSolver requires a regularly spaced grid")
        super().__init__(grid)


class FourierTridiagonalPoissonSolver(_PoissonSolver):
    def __init__(self, grid):
        if grid.regular[2]:
            raise ValueError("FourierTridiagonalPoissonSolver requires a stretched z")
        super().__init__(grid)


def solve(solver, rhs):
    """solve!(ϕ, solver, rhs): rhs is a real (Nz, Ny, Nx) numpy array (x fastest); returns ϕ of the same shape.
    For the Fourier-tridiagonal solver the multiplication by Δzᶜ (set_source_term!) is applied here on the host,
    exactly like `set_source_term!` does before the transforms."""
    g = solver.grid
    rhs = np.asarray(rhs, dtype=g.FT).reshape(g.N[::-1])
    if isinstance(solver, FourierTridiagonalPoissonSolver):
        H = g.H[2]
        dzc = g.dC[2][H:H + g.N[2]]  # Δzᶜ(k), k = 1..Nz  (index k + H - 1)
        rhs = (rhs * dzc[:, None, None]).astype(g.FT)
    solver._rhs.upload(rhs)
    _abi.call("ob_poisson_solve", solver.handle, solver._rhs.ptr, solver._phi.ptr)
    return solver._phi.download(g.N[::-1], g.FT)


class BatchedTridiagonalSolver:
    """BatchedTridiagonalSolver(grid; lower_diagonal, diagonal, upper_diagonal) along z with 1-D off-diagonals
    (length Nz-1) and a 3-D diagonal."""

    def __init__(self, grid, lower_diagonal, diagonal, upper_diagonal):
        self.grid, self.arch = grid, grid.architecture
        FT = grid.FT
        Nx, Ny, Nz = grid.N
        lo = np.ascontiguousarray(lower_diagonal, FT); up = np.ascontiguousarray(upper_diagonal, FT)
        dg = np.ascontiguousarray(np.broadcast_to(np.asarray(diagonal, FT), (Nz, Ny, Nx)))
        self._a = _DeviceBuffer(self.arch, lo.nbytes); self._a.upload(lo)
        self._c = _DeviceBuffer(self.arch, up.nbytes); self._c.upload(up)
        self._b = _DeviceBuffer(self.arch, dg.nbytes); self._b.upload(dg)
        self._t = _DeviceBuffer(self.arch, dg.nbytes)
        cn = Nx * Ny * Nz * 2 * np.dtype(FT).itemsize
        self._f = _DeviceBuffer(self.arch, cn); self._phi = _DeviceBuffer(self.arch, cn)

    def solve(self, rhs):
        g = self.grid
        FT = g.FT
        CT = np.complex128 if FT == np.float64 else np.complex64
        is_complex = np.iscomplexobj(rhs)
        f = np.ascontiguousarray(np.asarray(rhs).astype(CT).reshape(g.N[::-1]))
        self._f.upload(f)
        _abi.call("ob_batched_tridiagonal_solve", self.arch.ctx, _abi.OB_F64 if FT == np.float64 else _abi.OB_F32, 1,
                  g.N[0], g.N[1], g.N[2], self._a.ptr, self._b.ptr, self._c.ptr, self._f.ptr, self._phi.ptr, self._t.ptr)
        out = self._phi.download(g.N[::-1], CT)
        return out if is_complex else out.real.astype(FT)
