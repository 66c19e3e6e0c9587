"""CPU: DynamicSmagorinsky (directional averaging).  (1) the numpy oracle (oracle/dynsmag.py) against analytic answers;
(2) the kernels' own pointwise code (oceananigans.jl_b200/csrc/dynsmag.cuh, `__host__ __device__`) compiled for the host by
nvcc (tests/host_dynsmag.cu) against that oracle on random fields -- the arithmetic the GPU runs is checked without a GPU;
the GPU tests (test_gpu_parity.py: dynsmag_* configurations) then check launch geometry and plumbing."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import stretched_faces
from oracle import model as M
from oracle import dynsmag as DS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _model(size, topo, z=None, averaging=(1, 2), seed=0, amp=0.1):
    ext = ((0, 1.0), (0, 1.2), (-0.8, 0.0) if z is None else z)
    g = M.Grid(size, ext, topology=tuple(topo), halo=(3, 3, 3))
    om = M.Model(g, advection=("weno", 5), closure=[M.DynamicSmagorinsky(averaging=averaging)], tracers=("c",))
    rng = np.random.default_rng(seed)
    sh = lambda f: f.interior.shape
    om.set(u=amp * rng.standard_normal(sh(om.u)), v=amp * rng.standard_normal(sh(om.v)), w=amp * rng.standard_normal(sh(om.w)))
    om.update_state()
    return om


def test_laminar_shear_has_no_resolved_stress_and_a_known_MM():
    """u = S z: every filtered linear field is itself, so Lᵢⱼ Mᵢⱼ = 0 (L₁₁ = S² Δz²/6 meets M₁₁ = 0; M₁₃ ≠ 0 meets L₁₃ = 0) and
    MM = 2 M₁₃² with M₁₃ = 2 Δᶠ² (1 - ᾱ²β) Σ Σ₁₃, Σ = |S|/√2, Σ₁₃ = S/2, i.e. MM = 9 Δᶠ⁴ S⁴ / 4 ... evaluated exactly below."""
    S = 0.7
    g = M.Grid((6, 6, 16), ((0, 1.2), (0, 0.6), (-1.6, 0.0)), topology=("P", "P", "B"), halo=(3, 3, 3))
    om = M.Model(g, advection=("centered", 2), closure=[M.DynamicSmagorinsky(averaging=(1, 2))], tracers=("c",))
    zc = g.nodes(2, "c")[:, None, None]
    om.u.interior[...] = S * zc * np.ones((1, 6, 6))
    om.update_state()
    cf = om.dynamic_fields[0]
    d3 = 0.2 * 0.1 * 0.1
    Df2 = np.cbrt(d3) ** 2
    Sig, S13 = abs(S) / np.sqrt(2), S / 2
    M13 = 2 * Df2 * (1 - 4) * Sig * S13
    mm = cf["MM"].interior[4:-4]          # away from the walls (one-sided halo values there)
    assert np.allclose(mm, 2 * M13 ** 2, rtol=1e-10)
    assert np.abs(cf["LM"].interior[4:-4]).max() < 1e-12 * (2 * M13 ** 2)
    assert np.allclose(cf["Sigma"].interior[3:-3], Sig, rtol=1e-12)
    assert om.nue[0].interior[4:-4].max() < 1e-25           # cˢ² = minimum_numerator / MM


def test_coefficient_follows_the_averaging_dimensions():
    """sizes of 𝒥ᴸᴹ, 𝒥ᴹᴹ for averaging = 1, (1, 2), (2, 3), Colon (test/test_turbulence_closures.jl:421-448; numpy order (k, j, i))"""
    for dims, shape in (((1, 2), (10, 1, 1)), ((1, 2, 3), (1, 1, 1)), ((1,), (10, 6, 1)), ((3,), (1, 6, 8)), ((2, 3), (1, 1, 8))):
        om = _model((8, 6, 10), "PPB", averaging=dims)
        cf = om.dynamic_fields[0]
        assert cf["JLM"].shape == shape and cf["JMM"].shape == shape
        axes = tuple(2 - (d - 1) for d in dims)
        assert np.allclose(cf["JMM"], cf["MM"].interior.mean(axis=axes, keepdims=True), rtol=1e-13)


# ---- the kernels' pointwise code on the host ------------------------------------------------------------------------------
class HostField(C.Structure):
    _fields_ = [("p", C.c_void_p), ("Px", C.c_int), ("Py", C.c_int)]


class HostArgs(C.Structure):
    _fields_ = [("N", C.c_int * 3), ("H", C.c_int * 3), ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("dzc", C.c_void_p), ("rdzc", C.c_void_p), ("rdzf", C.c_void_p), ("koff", C.c_int)] + \
               [(n, HostField) for n in ("u", "v", "w", "ub", "vb", "wb", "Sg", "Sb", "LM", "MM", "nue")] + \
               [("J", C.c_void_p), ("avg", C.c_int * 3), ("JLM_min", C.c_double)]


@pytest.fixture(scope="module")
def hostlib():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "oracle", "_build", "libdynsmag_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "host_dynsmag.cu")
    hdr = os.path.join(ROOT, "oceananigans.jl_b200", "csrc", "dynsmag.cuh")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        # host code only is exercised; -fmad=false: no contraction on the device side either (irrelevant here), and the host
        # compiler is told not to contract so that the comparison is about the formulas
        subprocess.run([nvcc if os.path.exists(nvcc) else "nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
                        "-gencode", "arch=compute_100a,code=sm_100a", src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.dynsmag_host.restype = C.c_int
    return lib


def _hf(arr):
    f = HostField()
    f.p = arr.ctypes.data
    f.Px, f.Py = arr.shape[2], arr.shape[1]
    return f


@pytest.mark.parametrize("topo,stretched,dims", [("PPP", False, (1, 2)), ("PPB", False, (1, 2)), ("PPB", True, (1, 2)), ("BBB", True, (1, 2, 3)),
                                                 ("PBP", False, (1,))])
def test_kernel_arithmetic_on_the_host_matches_the_numpy_oracle(hostlib, topo, stretched, dims):
    size = (9, 8, 10)
    z = stretched_faces(size[2], 0.8) if stretched else None
    om = _model(size, topo, z=z, averaging=dims, seed=3)
    g = om.grid
    cf = om.dynamic_fields[0]
    a = HostArgs()
    for d in range(3):
        a.N[d], a.H[d] = g.N[d], g.H[d]
        a.avg[d] = int((d + 1) in dims)
    a.dx, a.dy = float(g.dC(0, np.array([1]))[0]), float(g.dC(1, np.array([1]))[0])
    ks = np.arange(1 - g.H[2], g.N[2] + g.H[2] + 1)
    dzc = np.ascontiguousarray(g.dC(2, ks), dtype=np.float64)
    dzf = np.ascontiguousarray(g.dF(2, ks), dtype=np.float64)
    rdzc, rdzf = 1.0 / dzc, 1.0 / dzf
    a.dz = float(dzc[g.H[2]])
    if stretched:
        a.dzc, a.rdzc, a.rdzf, a.koff = dzc.ctypes.data, rdzc.ctypes.data, rdzf.ctypes.data, g.H[2] - 1
    a.JLM_min = 1e-32
    work = {n: np.zeros_like(om.u.data if n == "ub" else om.v.data if n == "vb" else om.w.data if n == "wb" else cf["Sigma"].data)
            for n in ("ub", "vb", "wb", "Sg", "Sb", "LM", "MM", "nue")}
    a.u, a.v, a.w = _hf(om.u.data), _hf(om.v.data), _hf(om.w.data)
    for n, arr in work.items():
        setattr(a, n, _hf(arr))
    hostlib.dynsmag_host(1, C.byref(a))
    hostlib.dynsmag_host(2, C.byref(a))
    H, N = g.H, g.N
    inner = (slice(H[2], H[2] + N[2]), slice(H[1], H[1] + N[1]), slice(H[0], H[0] + N[0]))
    assert np.allclose(work["Sg"][inner], cf["Sigma"].interior, rtol=1e-13, atol=1e-300)
    assert np.allclose(work["Sb"][inner], cf["Sigmabar"].interior, rtol=1e-12, atol=1e-300)
    # halos of Σ, Σ̄ as the library fills them between the two kernels: the oracle's own fill on the same numbers
    for n, key in (("Sg", "Sigma"), ("Sb", "Sigmabar")):
        f = M.Field(g, "ccc", None, n)
        f.data[...] = work[n]
        M.fill_halo_regions(f)
        work[n][...] = f.data
    hostlib.dynsmag_host(3, C.byref(a))
    scale = np.abs(cf["MM"].interior).max()
    assert np.abs(work["MM"][inner] - cf["MM"].interior).max() <= 1e-11 * scale
    assert np.abs(work["LM"][inner] - cf["LM"].interior).max() <= 1e-11 * np.abs(cf["LM"].interior).max()
    axes = tuple(2 - (d - 1) for d in dims)
    J = np.concatenate([work["LM"][inner].mean(axis=axes).ravel(), work["MM"][inner].mean(axis=axes).ravel()]).astype(np.float64)
    a.J = J.ctypes.data
    hostlib.dynsmag_host(4, C.byref(a))
    want = om.nue[0].interior
    assert np.abs(work["nue"][inner] - want).max() <= 1e-10 * max(np.abs(want).max(), 1e-300)
    assert np.abs(want).max() > 0
