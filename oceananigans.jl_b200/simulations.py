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
    def __init__(self, interval):
        self.interval, self.next = float(interval), float(interval)

    def __call__(self, model):
        if model.clock.time >= self.next - 1e-12 * max(1.0, abs(self.next)):
            self.next += self.interval
            return True
        return False


class Callback:
    def __init__(self, func, schedule=None):
        self.func, self.schedule = func, schedule or IterationInterval(1)


class NaNChecker:
    """Diagnostics/nan_checker.jl: `any(isnan, parent(field))` on the first prognostic field (u); stops the run."""

    def __init__(self, fields):
        self.fields = fields

    def __call__(self, sim):
        for name, f in self.fields.items():
            if f.any_nan():
                sim.running = False
                raise FloatingPointError("time = %s, iteration = %d: NaN found in field %s. Aborting simulation."
                                         % (sim.model.clock.time, sim.model.clock.iteration, name))


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
        path = "%s_iteration%d.npz" % (self.prefix, model.clock.iteration)
        np.savez(path, **out)
        self.written.append(path)
        return path


class Checkpointer:
    """Checkpointer(model, schedule; prefix): writes restartable states (checkpoint.py); `restore(model, path)` resumes"""

    def __init__(self, model, schedule, prefix="checkpoint"):
        self.schedule, self.prefix, self.written = schedule, prefix, []

    def write(self, model):
        from .checkpoint import checkpoint
        path = checkpoint(model, "%s_iteration%d.npz" % (self.prefix, model.clock.iteration))
        self.written.append(path)
        return path


def conjure_time_step_wizard(sim, schedule=None, **kw):
    """conjure_time_step_wizard!(sim, IterationInterval(10); cfl=...)"""
    sim.callbacks["time_step_wizard"] = Callback(TimeStepWizard(**kw), schedule or IterationInterval(10))


def _aligned_time_step(sim, dt):
    """run.jl:40-58: do not step past stop_time"""
    clk = sim.model.clock
    if math.isfinite(sim.stop_time):
        remaining = sim.stop_time - clk.time
        if remaining > 0:
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
