"""GPU component tests through the C ABI: halo fills (bit-exact), Poisson solvers, batched tridiagonal solver,
diagnostics, error behaviour, Simulation driver."""
import numpy as np
import pytest

from helpers import Config, rel_l2, stretched_faces
from oracle import model as M

pytestmark = pytest.mark.gpu


def _laplacian_residual(og, phi, rhs):
    """∇²ϕ computed with the oracle grid's metrics after an oracle halo fill (test_poisson_solvers.jl:115-135)"""
    f = M.Field(og, "ccc")
    f.interior[...] = phi
    M.fill_halo_regions(f)
    Nx, Ny, Nz = og.N
    ft = og.ft
    c = lambda di, dj, dk: f.view(1 + di, Nx + di, 1 + dj, Ny + dj, 1 + dk, Nz + dk)
    lap = np.zeros_like(phi, dtype=np.float64)
    i = np.arange(1, Nx + 1); j = np.arange(1, Ny + 1); k = np.arange(1, Nz + 1)
    if og.topo[0] != M.FLAT:
        dxf = og.dF(0, i)[None, None, :]; dxfp = og.dF(0, i + 1)[None, None, :]; dxc = og.dC(0, i)[None, None, :]
        lap += ((c(1, 0, 0) - c(0, 0, 0)) / dxfp - (c(0, 0, 0) - c(-1, 0, 0)) / dxf) / dxc
    if og.topo[1] != M.FLAT:
        dyf = og.dF(1, j)[None, :, None]; dyfp = og.dF(1, j + 1)[None, :, None]; dyc = og.dC(1, j)[None, :, None]
        lap += ((c(0, 1, 0) - c(0, 0, 0)) / dyfp - (c(0, 0, 0) - c(0, -1, 0)) / dyf) / dyc
    if og.topo[2] != M.FLAT:
        dzf = og.dF(2, k)[:, None, None]; dzfp = og.dF(2, k + 1)[:, None, None]; dzc = og.dC(2, k)[:, None, None]
        lap += ((c(0, 0, 1) - c(0, 0, 0)) / dzfp - (c(0, 0, 0) - c(0, 0, -1)) / dzf) / dzc
    return lap


@pytest.mark.parametrize("topology", ["PPP", "PPB", "PBP", "BPP", "PBB", "BPB", "BBP", "BBB", "PPF", "PFB", "FBB"])
@pytest.mark.parametrize("size", [(16, 16, 16), (7, 11, 16), (11, 7, 13)])
def test_fft_poisson_solver(arch, topology, size):
    """∇²ϕ == rhs for a random mean-free rhs on every topology, incl. prime sizes and Flat dimensions
    (test/test_poisson_solvers.jl:68-106), and ϕ equals the oracle's FFTW-equivalent solve to 1e-11"""
    import ocean_b200 as ob
    size = tuple(1 if t == "F" else n for n, t in zip(size, topology))
    cfg = Config(size, tuple(None if t == "F" else (0, 1.0 + 0.5 * d) for d, t in enumerate(topology)), topology)
    og = cfg.oracle_grid()
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal(size[::-1])
    rhs -= rhs.mean()
    solver = ob.FFTBasedPoissonSolver(cfg.b200_grid(arch))
    phi = ob.solve(solver, rhs)
    ref = M.FFTPoissonSolver(og).solve(rhs)
    assert rel_l2(phi, ref) <= 1e-11
    lap = _laplacian_residual(og, phi, rhs)
    assert np.max(np.abs(lap - rhs)) <= 1e-9 * max(1.0, np.max(np.abs(rhs)))
    # the (1,1,1) mode is zeroed in every topology (fft_based_poisson_solver.jl:110): the solution is mean-free
    assert abs(phi.mean()) <= 1e-12 * max(1.0, np.abs(phi).max())


@pytest.mark.parametrize("topology", ["PPB", "PBB", "BPB", "BBB", "FPB"])
@pytest.mark.parametrize("ft", [np.float64, np.float32])
def test_fourier_tridiagonal_solver(arch, topology, ft):
    import ocean_b200 as ob
    size = tuple(1 if t == "F" else n for n, t in zip((12, 10, 14), topology))
    ext = [None if t == "F" else (0, 2.0) for t in topology]
    ext[2] = stretched_faces(14, 3.0)
    cfg = Config(size, tuple(ext), topology, ft=ft)
    og = cfg.oracle_grid()
    rng = np.random.default_rng(6)
    rhs = rng.standard_normal(size[::-1]).astype(ft)
    k = np.arange(1, size[2] + 1)
    dzc = og.dC(2, k)[:, None, None]
    rhs = (rhs - (rhs * dzc).sum() / (dzc.sum() * size[0] * size[1])).astype(ft)  # volume-mean free
    solver = ob.FourierTridiagonalPoissonSolver(cfg.b200_grid(arch))
    phi = ob.solve(solver, rhs)
    ref = M.FourierTridiagonalPoissonSolver(og).solve((rhs * dzc).astype(ft))
    tol = 1e-11 if ft == np.float64 else 2e-4
    assert rel_l2(phi, ref) <= tol
    if ft == np.float64:
        lap = _laplacian_residual(og, phi, rhs)
        assert np.max(np.abs(lap - rhs)) <= 1e-8 * max(1.0, np.max(np.abs(rhs)))


def test_batched_tridiagonal_solver_vs_dense(arch):
    """test/test_batched_tridiagonal_solver.jl:8-45: against a dense solve, real and complex right-hand sides"""
    import ocean_b200 as ob
    cfg = Config((4, 3, 9), ((0, 1.0),) * 3, "PPB")
    grid = cfg.b200_grid(arch)
    rng = np.random.default_rng(7)
    Nz = 9
    a, c = rng.random(Nz - 1), rng.random(Nz - 1)
    b = 3 + rng.random((Nz, 3, 4))
    f = rng.standard_normal((Nz, 3, 4)) + 1j * rng.standard_normal((Nz, 3, 4))
    solver = ob.BatchedTridiagonalSolver(grid, lower_diagonal=a, diagonal=b, upper_diagonal=c)
    phi = solver.solve(f)
    for j in range(3):
        for i in range(4):
            A = np.diag(b[:, j, i]) + np.diag(a, -1) + np.diag(c, 1)
            assert np.allclose(phi[:, j, i], np.linalg.solve(A, f[:, j, i]), rtol=1e-12, atol=1e-13)
    phir = solver.solve(f.real)
    assert np.allclose(phir, phi.real, rtol=1e-13)


HALO_CASES = [
    ("PPP", (3, 3, 3), {}),
    ("PPB", (3, 3, 3), {"bottom": ("Value", 0.3), "top": ("Gradient", -0.7)}),
    ("BBB", (2, 3, 4), {"west": ("Flux", 1.0), "east": ("Value", 2.0), "south": ("Gradient", 0.1), "north": ("Flux", None),
                        "bottom": ("Value", -1.0), "top": ("Flux", 3.0)}),
    ("PBF", (3, 2, 0), {"south": ("Value", 1.5)}),
    ("BPP", (1, 1, 1), {}),
]


@pytest.mark.parametrize("ft", [np.float64, np.float32])
@pytest.mark.parametrize("case", range(len(HALO_CASES)))
def test_halo_fill_bit_exact(arch, case, ft):
    """fill_halo_regions! on random parents (stale halos included) is bit-identical to the reference ordering and
    kernels for every field location (test/test_halo_regions.jl:22-41 extended to Value/Gradient/Flux BCs)"""
    import ocean_b200 as ob
    topology, halo, tracer_bcs = HALO_CASES[case]
    size = tuple(1 if t == "F" else n for n, t in zip((9, 8, 7), topology))
    ext = tuple(None if t == "F" else (0, 1.0 + d) for d, t in enumerate(topology))
    if topology[2] == "B":
        ext = ext[:2] + (stretched_faces(size[2], 2.0),)
    cfg = Config(size, ext, topology, halo=halo, ft=ft, advection=("centered", 2), tracers=("c",), bcs={"c": tracer_bcs})
    om = cfg.oracle_model()
    bm = cfg.b200_model(arch)
    rng = np.random.default_rng(100 + case)
    for name, of in (("u", om.u), ("v", om.v), ("w", om.w), ("c", om.tracers[0]), ("pNHS", om.pNHS)):
        bf = {**bm.velocities, **bm.tracers, **bm.pressures}[name]
        for fill_normal in (True, False):
            parent = rng.standard_normal(of.data.shape).astype(ft)
            of.data[...] = parent
            bf.set_parent(parent)
            M.fill_halo_regions(of, fill_normal_flow_bcs=fill_normal)
            bm.fill_halo_regions(name, fill_normal_flow_bcs=fill_normal)
            got = bf.parent()
            assert np.array_equal(got.view(np.uint8), of.data.view(np.uint8)), (name, fill_normal)
            # the model-free entry point (ob_fill_halo_array: any Field, as the reference's fill_halo_regions!) is the same fill
            bf.set_parent(parent)
            bf.fill_halo_regions(fill_normal_flow_bcs=fill_normal)
            assert np.array_equal(bf.parent().view(np.uint8), of.data.view(np.uint8)), (name, fill_normal, "ob_fill_halo_array")


def test_advection_timescale_and_nan_checker(arch):
    import ocean_b200 as ob
    cfg = Config((12, 10, 8), ((0, 1.0), (0, 2.0), stretched_faces(8, 1.0)), "PPB", advection=("centered", 2))
    bm = cfg.b200_model(arch)
    ic = cfg.initial_conditions(3)
    ob.set(bm, **ic)
    u, v, w = (bm.velocities[n].interior().astype(np.float64) for n in "uvw")
    g = bm.grid
    H = g.H[2]
    dzf = g.dF[2][H + 1:H + 1 + g.N[2]].astype(np.float64)  # Δzᶠ(k), k = 1..Nz
    inv = np.abs(u) / float(g.dF[0]) + np.abs(v) / float(g.dF[1]) + np.abs(w[:g.N[2]]) / dzf[:, None, None]
    assert np.isclose(bm.cell_advection_timescale(), (1 / inv).min(), rtol=1e-12)
    assert not bm.velocities["u"].any_nan()
    p = bm.velocities["u"].parent()
    p[2, 3, 4] = np.nan
    bm.velocities["u"].set_parent(p)
    assert bm.velocities["u"].any_nan()


def test_unsupported_options_raise(arch):
    import ocean_b200 as ob
    grid = ob.RectilinearGrid(arch, size=(8, 8, 8), extent=(1, 1, 1))
    with pytest.raises(ob.OceanB200Error):
        ob.NonhydrostaticModel(grid, forcing={"u": lambda x, y, z, t: 0.0})
    with pytest.raises(ob.OceanB200Error):
        ob.NonhydrostaticModel(grid, tracers=("c",), boundary_conditions={"c": ob.FieldBoundaryConditions(top=ob.FluxBoundaryCondition(lambda x, y, t: 1.0))})
    with pytest.raises(ob.OceanB200Error):
        ob.NonhydrostaticModel(grid, advection=ob.WENO(order=13))
    with pytest.raises(NotImplementedError):
        ob.CPU()


def test_simulation_run_taylor_green(arch):
    """run!(Simulation) on the Taylor-Green vortex (test/test_dynamics.jl:214-259): 64x64x2, ν = 1, rel. error < 5e-6"""
    import ocean_b200 as ob
    for ts in ("RungeKutta3", "QuasiAdamsBashforth2"):
        grid = ob.RectilinearGrid(arch, size=(64, 64, 2), x=(0, 2 * np.pi), y=(0, 2 * np.pi), z=(0, 2 * np.pi),
                                  topology=(ob.Periodic, ob.Periodic, ob.Bounded))
        model = ob.NonhydrostaticModel(grid, timestepper=ts, closure=ob.ScalarDiffusivity(nu=1.0))
        ob.set(model, u=lambda x, y, z: np.cos(x) * np.sin(y), v=lambda x, y, z: -np.sin(x) * np.cos(y))
        dt = 1e-4
        sim = ob.Simulation(model, Δt=dt, stop_iteration=10)
        ob.run(sim)
        assert model.clock.iteration == 10
        t = model.clock.time
        xu, yu, _ = model.velocities["u"].nodes()
        u = model.velocities["u"].interior()
        ua = np.exp(-2 * t) * np.cos(xu)[None, None, :] * np.sin(yu)[None, :, None] * np.ones((2, 1, 1))
        assert np.abs(u - ua).max() / np.abs(ua).max() < 5e-6


def test_time_step_wizard(arch):
    import ocean_b200 as ob
    grid = ob.RectilinearGrid(arch, size=(16, 16, 16), extent=(1, 1, 1))
    model = ob.NonhydrostaticModel(grid, advection=ob.WENO())
    rng = np.random.default_rng(1)
    ob.set(model, u=rng.uniform(-1, 1, (16, 16, 16)), v=rng.uniform(-1, 1, (16, 16, 16)))
    sim = ob.Simulation(model, Δt=1e-4, stop_iteration=4)
    ob.conjure_time_step_wizard(sim, ob.IterationInterval(2), cfl=0.5, max_change=1.5)
    ob.run(sim)
    assert 1e-4 < sim.dt <= 1e-4 * 1.5 ** 3


def test_host_streamed_stepper_matches_resident_stepping(arch):
    """ocean_b200.HostStreamedStepper: three members living in pinned host memory, each advanced through its own device
    lane (upload, time_step!, asynchronous download), reproduce resident time stepping bit for bit"""
    import ctypes as C
    import ocean_b200 as ob
    cfg = Config((24, 20, 16), ((0, 2 * np.pi),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)],
                 buoyancy=("tracer",), tracers=("b",))
    ics = [cfg.initial_conditions(seed) for seed in (1, 2, 3)]
    # resident reference: one model per member, 3 steps each
    want = []
    for ic in ics:
        m = cfg.b200_model(arch)
        ob.set(m, **ic)
        for _ in range(3):
            ob.time_step(m, 1e-3)
        want.append([f.parent() for f in m.prognostic_fields.values()])
        del m
    stepper = ob.HostStreamedStepper(lambda a: cfg.b200_model(a), lanes=3, device=arch.device)
    members = []
    for lane, ic in enumerate(ics):
        ob.set(stepper.models[lane], **ic)
        mem = stepper.new_member()
        stepper.download(mem, lane)
        members.append(mem)
    for s in range(9):
        mem = members[s % 3]
        assert stepper.step(mem, mem, 1e-3) == s % 3
    stepper.join_into(0)
    stepper.synchronize()
    for mem, ref, model in zip(members, want, stepper.models):
        for p, nb, r, f in zip(mem.ptrs, mem.nbytes, ref, model.prognostic_fields.values()):
            got = np.frombuffer((C.c_char * nb).from_address(p.value), dtype=r.dtype).reshape(r.shape)
            assert np.array_equal(got, r), f.name
        mem.free()


def test_host_streamed_stepper_with_more_members_than_lanes(arch):
    """four members through TWO lanes: a lane then receives a member other than the one it held last, whose tendencies and
    clock must not leak into it (update_state! is re-run after the upload; the member's own clock is installed)"""
    import ctypes as C
    import ocean_b200 as ob
    cfg = Config((24, 20, 16), ((0, 2 * np.pi),) * 3, "PPP", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)],
                 buoyancy=("tracer",), tracers=("b",))
    ics = [cfg.initial_conditions(seed) for seed in (1, 2, 3, 4)]
    want = []
    for ic in ics:
        m = cfg.b200_model(arch)
        ob.set(m, **ic)
        for _ in range(3):
            ob.time_step(m, 1e-3)
        want.append([f.parent() for f in m.prognostic_fields.values()])
        del m
    stepper = ob.HostStreamedStepper(lambda a: cfg.b200_model(a), lanes=2, device=arch.device)
    members = []
    for n, ic in enumerate(ics):
        ob.set(stepper.models[n % 2], **ic)
        mem = stepper.new_member()
        stepper.download(mem, n % 2)
        mem.clock = (0.0, 0, 0.0, 0.0)
        members.append(mem)
    for s in range(12):
        mem = members[s % 4]
        assert stepper.step(mem, mem, 1e-3) == (s % 4) % 2   # a member stays on the lane of its first step
    stepper.synchronize()
    for mem, ref in zip(members, want):
        assert mem.clock[1] == 3
        for p, nb, r in zip(mem.ptrs, mem.nbytes, ref):
            got = np.frombuffer((C.c_char * nb).from_address(p.value), dtype=r.dtype).reshape(r.shape)
            assert np.array_equal(got, r)
        mem.free()


@pytest.mark.parametrize("ts", ["rk3", "ab2"])
def test_checkpoint_restore_continues_bit_identically(arch, ts, tmp_path):
    """checkpoint (parents of the prognostic fields, G⁻, clock) -> fresh model -> restore -> same trajectory, bit for bit
    (checkpointer.jl; test_checkpointer.jl compares a restored run with an uninterrupted one)"""
    import ocean_b200 as ob
    cfg = Config((16, 12, 10), ((0, 1.6), (0, 1.2), (-1.0, 0.0)), "PPB", advection=("weno", 5), closure=[("lilly", 0.16, 1.0, 1.0)],
                 buoyancy=("tracer",), coriolis_f=0.2, tracers=("b", "c"), timestepper=ts, bcs={"b": {"top": ("Flux", 1e-4)}})
    a = cfg.b200_model(arch)
    ob.set(a, **cfg.initial_conditions(12))
    for _ in range(3):
        ob.time_step(a, 1e-2)
    path = ob.checkpoint(a, str(tmp_path / "ckpt.npz"))
    for _ in range(3):
        ob.time_step(a, 1e-2)
    b = cfg.b200_model(arch)
    ob.restore(b, path)
    assert b.clock.iteration == 3 and b.clock.time == pytest.approx(3e-2)
    for _ in range(3):
        ob.time_step(b, 1e-2)
    for (n, fa), fb in zip(a.prognostic_fields.items(), b.prognostic_fields.values()):
        assert np.array_equal(fa.parent(), fb.parent()), n


def test_simulation_output_writer_and_checkpointer(arch, tmp_path):
    """run!(simulation) with an output writer and a checkpointer (host plumbing: device -> host copies on a schedule)"""
    import ocean_b200 as ob
    cfg = Config((16, 16, 8), ((0, 1.0), (0, 1.0), (-0.5, 0.0)), "PPB", advection=("weno", 5), closure=[("scalar", 1e-3, 1e-3)],
                 buoyancy=("tracer",), tracers=("b",))
    m = cfg.b200_model(arch)
    ob.set(m, **cfg.initial_conditions(2))
    sim = ob.Simulation(m, Δt=1e-3, stop_iteration=6)
    sim.output_writers["fields"] = ob.NPZOutputWriter(m, {"u": m.velocities["u"], "b": m.tracers["b"]}, ob.IterationInterval(3),
                                                      prefix=str(tmp_path / "out"))
    sim.output_writers["ckpt"] = ob.Checkpointer(m, ob.IterationInterval(6), prefix=str(tmp_path / "ckpt"))
    ob.run(sim)
    w = sim.output_writers["fields"].written
    assert [p.split("iteration")[-1] for p in w] == ["0.npz", "3.npz", "6.npz"]
    z = np.load(w[-1])
    assert z["u"].shape == (8, 16, 16) and int(z["iteration"]) == 6 and np.array_equal(z["b"], m.tracers["b"].interior())
    m2 = cfg.b200_model(arch)
    ob.restore(m2, sim.output_writers["ckpt"].written[-1])
    assert m2.clock.iteration == 6 and np.array_equal(m2.velocities["u"].parent(), m.velocities["u"].parent())
