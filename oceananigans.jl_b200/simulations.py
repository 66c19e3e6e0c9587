"""Host-side mirror of `Simulation(model; Δt, stop_iteration, stop_time)` and `run!` -- unchanged driver logic on top
of the B200 `time_step!`.

Reference: src/Simulations/simulation.jl:12-119 (struct, default NaNChecker every 100 iterations :91-94),
run.jl:40-58 (aligned_time_step), :151-187 (run!), :216-256 (time_step!(sim)), time_step_wizard.jl
(TimeStepWizard / new_time_step), src/Diagnostics/nan_checker.jl.  All of it is host logic; the device work is
`time_step!(model, Δt)`, `ob_cell_advection_timescale` and `ob_any_nan`.
"""
import math
import time as _time

from . import models as _models


class IterationInterval:
    def __init__(self, interval, offset=0):
        self.interval, self.offset = int(interval), int(offset)

    def __call__(self, model):
        return (model.clock.iteration - self.offset) % self.interval == 0


class TimeInterval:
    """Utils/schedules.jl: fires at first_actuation_time + n * interval; the first actuation time is the clock time when the
    simulation is initialised (so a writer fires at iteration 0, and after a restore at t > 0 it does not fire on every
    iteration until it has caught up); next = first + (actuations + 1) * interval, without accumulated rounding."""

    def __init__(self, interval):
        self.interval = float(interval)
        self.first, self.actuations = None, 0

    def initialize(self, model):
        self.first = float(model.clock.time)
        self.actuations = 0
        return True   # initialize!(schedule, model) actuates at initialisation

    def next_actuation_time(self):
        first = 0.0 if self.first is None else self.first
        return first + (self.actuations + 1) * self.interval

    def __call__(self, model):
        if self.first is None:
            return self.initialize(model)
        t = self.next_actuation_time()
        if model.clock.time >= t - 1e-12 * max(1.0, abs(t)):
            # catch up if several intervals were skipped by one long step
            while model.clock.time >= self.next_actuation_time() - 1e-12 * max(1.0, abs(t)):
                self.actuations += 1
            return True
        return False


class Callback:
    def __init__(self, func, schedule=None):
        self.func, self.schedule = func, schedule or IterationInterval(1)


class NaNChecker:
    """Diagnostics/nan_checker.jl: `any(isnan, parent(field))` on the first prognostic field (u).  Default erroring=False,
    as in the reference: the run is stopped (sim.running = False) and run! returns normally, so the final synchronize and the
    remaining writers still happen; erroring=True raises."""

    def __init__(self, fields, erroring=False):
        self.fields, self.erroring = fields, erroring
        self.message = None

    def __call__(self, sim):
        for name, f in self.fields.items():
            if f.any_nan():
                sim.running = False
                self.message = ("time = %s, iteration = %d: NaN found in field %s. Stopping simulation."
                                % (sim.model.clock.time, sim.model.clock.iteration, name))
                if self.erroring:
                    raise FloatingPointError(self.message)
                print("[NaNChecker] " + self.message)
                return


class TimeStepWizard:
    """TimeStepWizard(; cfl=0.2, max_change=1.1, min_change=0.5, max_Δt=Inf, min_Δt=0) (time_step_wizard.jl)."""

    def __init__(self, cfl=0.2, max_change=1.1, min_change=0.5, max_dt=math.inf, min_dt=0.0):
        self.cfl, self.max_change, self.min_change, self.max_dt, self.min_dt = cfl, max_change, min_change, max_dt, min_dt

    def new_time_step(self, old_dt, model):
        tau = model.cell_advection_timescale()  # min over the grid, on device
        new_dt = self.cfl * tau
        new_dt = min(self.max_change * old_dt, new_dt)
        new_dt = max(self.min_change * old_dt, new_dt)
        new_dt = max(self.min_dt, min(self.max_dt, new_dt))
        return new_dt

    def __call__(self, sim):
        sim.dt = self.new_time_step(sim.dt, sim.model)


class Simulation:
    def __init__(self, model, dt=None, stop_iteration=math.inf, stop_time=math.inf, wall_time_limit=math.inf, Δt=None, verbose=False):
        self.model = model
        self.dt = Δt if Δt is not None else dt
        if self.dt is None:
            raise ValueError("Simulation needs Δt")
        self.stop_iteration, self.stop_time, self.wall_time_limit = stop_iteration, stop_time, wall_time_limit
        self.callbacks = {}
        self.callbacks["nan_checker"] = Callback(NaNChecker({"u": model.velocities["u"]}), IterationInterval(100))
        self.output_writers = {}   # name -> NPZOutputWriter / Checkpointer (simulation.output_writers[:name] = ...)
        self.running = False
        self.initialized = False
        self.run_wall_time = 0.0
        self.verbose = verbose

    def add_callback(self, name, func, schedule=None):
        self.callbacks[name] = Callback(func, schedule)


class NPZOutputWriter:
    """Host-side output writer: the interiors of the named fields, copied device -> host when `schedule` fires and written
    as <prefix>_iteration<N>.npz with the clock (the role of JLD2Writer / NetCDFWriter, src/OutputWriters/: every writer
    starts with fetch_output's on_architecture(CPU(), ...) copy, fetch_output.jl:22).  `fields`: dict name -> Field."""

    def __init__(self, model, fields, schedule, prefix="output", with_halos=False):
        self.fields, self.schedule, self.prefix, self.with_halos = dict(fields), schedule, prefix, with_halos
        self.written = []

    def write(self, model):
        import numpy as np
        out = {n: (f.parent() if self.with_halos else np.ascontiguousarray(f.interior())) for n, f in self.fields.items()}
        out["time"], out["iteration"] = np.float64(model.clock.time), np.int64(model.clock.iteration)
        path = "%s%s_iteration%d.npz" % (self.prefix, _rank_suffix(model), model.clock.iteration)
        np.savez(path, **out)
        self.written.append(path)
        return path


def _rank_suffix(model):
    """"_rank{r}" under a distributed architecture: every rank writes its own x-slab (output_writer_utils.jl:266-267,
    checkpointer.jl:42); empty on one device"""
    arch = model.grid.architecture
    return "_rank%d" % arch.rank if getattr(arch, "world", 1) > 1 else ""


class Checkpointer:
    """Checkpointer(model, schedule; prefix): writes restartable states (checkpoint.py); `restore(model, path)` resumes"""

    def __init__(self, model, schedule, prefix="checkpoint"):
        self.schedule, self.prefix, self.written = schedule, prefix, []

    def write(self, model):
        from .checkpoint import checkpoint
        path = checkpoint(model, "%s%s_iteration%d.npz" % (self.prefix, _rank_suffix(model), model.clock.iteration))
        self.written.append(path)
        return path


def conjure_time_step_wizard(sim, schedule=None, **kw):
    """conjure_time_step_wizard!(sim, IterationInterval(10); cfl=...)"""
    sim.callbacks["time_step_wizard"] = Callback(TimeStepWizard(**kw), schedule or IterationInterval(10))


def _aligned_time_step(sim, dt):
    """run.jl:40-58: do not step past stop_time, nor past the next actuation of a TimeInterval callback / writer
    (schedule_aligned_time_step)"""
    clk = sim.model.clock
    if math.isfinite(sim.stop_time):
        remaining = sim.stop_time - clk.time
        if remaining > 0:
            dt = min(dt, remaining)
    for item in list(sim.callbacks.values()) + list(sim.output_writers.values()):
        sch = getattr(item, "schedule", None)
        if isinstance(sch, TimeInterval) and sch.first is not None:
            remaining = sch.next_actuation_time() - clk.time
            if remaining > 1e-12 * max(1.0, abs(clk.time)):
                dt = min(dt, remaining)
    return dt


def _stop_criteria(sim):
    clk = sim.model.clock
    if clk.iteration >= sim.stop_iteration:
        return True
    if clk.time >= sim.stop_time:
        return True
    if sim.run_wall_time >= sim.wall_time_limit:
        return True
    return False


def time_step_simulation(sim):
    """time_step!(sim::Simulation) (run.jl:216-256)"""
    t0 = _time.time()
    if not sim.initialized:
        sim.model.update_state()
        sim.initialized = True
        for cb in sim.callbacks.values():
            if cb.schedule(sim.model):
                cb.func(sim)
        for w in sim.output_writers.values():   # initialize!(sim): writers fire at iteration 0 too (run.jl:258-300)
            if w.schedule(sim.model):
                w.write(sim.model)
    dt = _aligned_time_step(sim, sim.dt)
    _models.time_step(sim.model, dt)
    for cb in sim.callbacks.values():
        if cb.schedule(sim.model):
            cb.func(sim)
    for w in sim.output_writers.values():
        if w.schedule(sim.model):
            w.write(sim.model)
    sim.run_wall_time += _time.time() - t0


def run(sim):
    """run!(sim) (run.jl:151-187)"""
    sim.running = True
    sim.run_wall_time = 0.0
    while sim.running and not _stop_criteria(sim):
        time_step_simulation(sim)
    sim.model.synchronize()
    sim.running = False
