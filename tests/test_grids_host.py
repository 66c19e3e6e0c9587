"""CPU: grid generation of the host mirror (oceananigans.jl_b200/grids.py) and of the oracle (oracle/model.py: Grid) against the
reference's own grid tests (test/test_grids.jl:59-150, 438-485), and against each other bit for bit -- they are two independent
restatements of grid_generation.jl:34-156 (exact rationals vs long double), and the spacings they produce are what crosses the
C ABI (ob_grid_desc)."""
import numpy as np
import pytest

from oracle import model as M


class _NoDevice:
    """RectilinearGrid only stores the architecture; descriptors and coordinates are host data"""
    ctx = None


def _product(size, topology, halo, ft=np.float64, **xyz):
    import ocean_b200 as ob
    T = {"P": ob.Periodic, "B": ob.Bounded, "F": ob.Flat}
    return ob.RectilinearGrid(_NoDevice(), ft, size=size, halo=halo, topology=tuple(T[t] for t in topology), **xyz)


def _oracle(size, topology, halo, ft=np.float64, **xyz):
    return M.Grid(size, (xyz["x"], xyz["y"], xyz["z"]), topology=tuple(topology), halo=halo, ft=ft)


class View:
    """logical (Julia OffsetArray) indexing of faces / centres / spacings for either implementation"""

    def __init__(self, g, kind):
        self.g, self.kind = g, kind
        self.H = g.H
        self.N = g.N

    def face(self, d, i):
        return self.g.faces[d][i - 1 + self.H[d]]

    def center(self, d, i):
        return self.g.centers[d][i - 1 + self.H[d]]

    def n_faces(self, d):
        return len(self.g.faces[d])

    def n_centers(self, d):
        return len(self.g.centers[d])

    def dF(self, d, i):   # Δᶠ (spacing located at faces) at logical i
        if self.kind == "oracle":
            return self.g.dF(d, np.asarray(i))
        a = self.g.dF[d]
        return a[np.asarray(i) + self.H[d]] if isinstance(a, np.ndarray) else np.full(np.shape(i), a)

    def dC(self, d, i):   # Δᶜ (cell widths) at logical i
        if self.kind == "oracle":
            return self.g.dC(d, np.asarray(i))
        a = self.g.dC[d]
        return a[np.asarray(i) + self.H[d] - 1] if isinstance(a, np.ndarray) else np.full(np.shape(i), a)


def make(kind, *a, **k):
    return View((_oracle if kind == "oracle" else _product)(*a, **k), kind)


KINDS = ["oracle", "product"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ft", [np.float64, np.float32])
def test_regular_rectilinear_halo_faces_first_cells_end_faces(kind, ft):
    """test_regular_rectilinear_correct_halo_faces / _first_cells / _end_faces (test_grids.jl:69-117)"""
    N, H, L = 4, 1, 2.0
    D = L / N
    g = make(kind, (N, N, N), "PBB", (H, H, H), ft, x=(0, L), y=(0, L), z=(0, L))
    for d in range(3):
        assert g.face(d, 0) == -H * D
    assert g.face(0, N + 1) == L                      # Periodic
    assert g.face(1, N + 2) == L + H * D and g.face(2, N + 2) == L + H * D
    g = make(kind, (N, N, N), "PPB", (H, H, H), ft, x=(0, 4.0), y=(0, 4.0), z=(0, 4.0))
    for d in range(3):
        assert g.center(d, 1) == 0.5


@pytest.mark.parametrize("kind", KINDS)
def test_regular_rectilinear_ranges_have_correct_length(kind):
    """test_regular_rectilinear_ranges_have_correct_length, _no_roundoff_error_in_ranges (test_grids.jl:119-149)"""
    Nx, Ny, Nz, Hx, Hy, Hz = 8, 9, 10, 1, 2, 1
    g = make(kind, (Nx, Ny, Nz), "BBB", (Hx, Hy, Hz), x=(0, 1), y=(0, 1), z=(0, 1))
    assert g.n_centers(0) == Nx + 2 * Hx and g.n_centers(1) == Ny + 2 * Hy and g.n_centers(2) == Nz + 2 * Hz
    assert g.n_faces(0) == Nx + 1 + 2 * Hx and g.n_faces(1) == Ny + 1 + 2 * Hy and g.n_faces(2) == Nz + 1 + 2 * Hz
    g = make(kind, (1, 1, 64), "PPB", (1, 1, 1), x=(0, 1), y=(0, 1), z=(-np.pi / 2, 0))   # issue #480
    assert g.n_centers(2) == 64 + 2 and g.n_faces(2) == 64 + 2 + 1


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("ft", [np.float64, np.float32])
def test_rectilinear_grid_correct_spacings(kind, ft):
    """test_rectilinear_grid_correct_spacings (test_grids.jl:438-485), the z part: tanh-stretched faces (x and y must be
    regular on this path: FFT / Fourier-tridiagonal solvers)"""
    N, S = 16, 3
    zf = lambda k: np.tanh(S * (2 * (k - 1) / N - 1)) / np.tanh(S)
    faces = np.array([zf(k) for k in range(1, N + 2)])
    g = make(kind, (N, N, N), "PPB", (3, 3, 3), ft, x=(0, N), y=(0, N), z=faces)
    k = np.arange(1, N + 1)
    assert np.all(g.dF(0, k) == 1) and np.all(g.dC(0, k) == 1)
    rt = 1e-6 if ft == np.float32 else 1e-12
    zc = lambda k: (zf(k) + zf(k + 1)) / 2
    assert np.allclose([g.face(2, q) for q in range(1, N + 2)], [zf(q) for q in range(1, N + 2)], rtol=rt)
    assert np.allclose([g.center(2, q) for q in k], [zc(q) for q in k], rtol=rt, atol=rt)
    assert np.allclose(g.dC(2, k), [zf(q + 1) - zf(q) for q in k], rtol=10 * rt, atol=rt)
    k2 = np.arange(2, N + 1)   # Δzᵃᵃᶠ[1] involves a halo point
    assert np.allclose(g.dF(2, k2), [zc(q) - zc(q - 1) for q in k2], rtol=10 * rt, atol=rt)


@pytest.mark.parametrize("ft", [np.float64, np.float32])
@pytest.mark.parametrize("topology,zext", [("PPB", (-0.7, 0.0)), ("PPP", (0.0, 2 * np.pi)), ("BBB", (-1.0, 1.0)), ("PPB", "stretched")])
def test_host_mirror_and_oracle_grids_agree_bit_for_bit(topology, zext, ft):
    """two independent restatements of grid_generation.jl produce identical nodes and spacings (what ob_grid_desc carries)"""
    N = (12, 10, 14)
    H = (3, 2, 3)
    z = -1.3 * (1 - np.tanh(1.1 * np.arange(15) / 14) / np.tanh(1.1)) if zext == "stretched" else zext
    kw = dict(x=(0.0, 2 * np.pi), y=(-1.0, 0.3), z=z)
    a, b = make("oracle", N, topology, H, ft, **kw), make("product", N, topology, H, ft, **kw)
    for d in range(3):
        assert np.array_equal(a.g.faces[d], b.g.faces[d]) and np.array_equal(a.g.centers[d], b.g.centers[d]), d
        idx = np.arange(1, N[d] + 1)
        assert np.array_equal(a.dF(d, idx), b.dF(d, idx)) and np.array_equal(a.dC(d, idx), b.dC(d, idx)), d
        assert a.g.L[d] == b.g.L[d]
