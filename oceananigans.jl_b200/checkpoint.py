"""Checkpoint / restore of a NonhydrostaticModel (host-side plumbing: SURVEY §8 row f1, §5 "checkpoint/resume").

Reference: src/OutputWriters/checkpointer.jl (prognostic fields + timestepper tendencies + clock written to JLD2;
`set!(model, filepath)` restores them) and the `prognostic_state` / `restore_prognostic_state!` pairs of the time
steppers.  Here the state is the parent arrays (halos included) of every prognostic field, G⁻ (needed by the RK3 second
and third stages only within a step, by AB2 across steps) and the clock; everything else (Gⁿ, pHY′, closure fields, pNHS)
is a function of that state and is rebuilt by `update_state!`.  A restored model continues bit for bit
(tests/test_gpu_components.py::test_checkpoint_restore_continues_bit_identically).
"""
import numpy as np


def checkpoint(model, path):
    """write the model state to `path` (.npz)"""
    out = {}
    for name, f in model.prognostic_fields.items():
        out["field__" + name] = f.parent()
    for n, f in enumerate(model.Gm):
        out["Gm__%d" % n] = f.parent()
    out["pNHS"] = model.pressures["pNHS"].parent()
    c = model.clock
    out["clock"] = np.array([c.time, c.iteration, c.stage, c.last_dt, c.last_stage_dt], dtype=np.float64)
    arch = model.grid.architecture
    out["meta"] = np.array([model.timestepper, str(np.dtype(model.grid.FT)), ",".join(model.tracer_names)])
    out["layout"] = np.array([getattr(arch, "rank", 0), getattr(arch, "world", 1)] + [int(n) for n in model.grid.N], dtype=np.int64)
    np.savez(path, **out)
    return path


def restore(model, path):
    """set!(model, checkpoint): `model` must have been built with the same grid, tracers and time stepper"""
    z = np.load(path, allow_pickle=False)
    ts, ft, names = [str(x) for x in z["meta"]]
    if ts != model.timestepper or names != ",".join(model.tracer_names) or ft != str(np.dtype(model.grid.FT)):
        raise ValueError("checkpoint %r was written by a different model (%s, %s, tracers %s)" % (path, ts, ft, names))
    if "layout" in z:
        arch = model.grid.architecture
        want = [getattr(arch, "rank", 0), getattr(arch, "world", 1)] + [int(n) for n in model.grid.N]
        if [int(x) for x in z["layout"]] != want:
            raise ValueError("checkpoint %r holds rank/world/local size %r, this model is %r" % (path, [int(x) for x in z["layout"]], want))
    for name, f in model.prognostic_fields.items():
        f.set_parent(z["field__" + name])
    for n, f in enumerate(model.Gm):
        f.set_parent(z["Gm__%d" % n])
    model.pressures["pNHS"].set_parent(z["pNHS"])
    c = model.clock
    t, it, st, ldt, lsdt = z["clock"]
    c.time, c.iteration, c.stage, c.last_dt, c.last_stage_dt = float(t), int(it), int(st), float(ldt), float(lsdt)
    model.update_state()
    return model
