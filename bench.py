#!/usr/bin/env python
"""bench.py -- cell-updates/s of the NonhydrostaticModel time step (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference, host cores

A "step" is one full `time_step!` (RK3: three stages, each with a pressure solve) of BASELINE.json configs[1]:
3-D triply-periodic NonhydrostaticModel, WENO-5, BuoyancyTracer, ScalarDiffusivity, Float64, 256^3 per GPU
(weak scaling: slab-x, 256*N x 256 x 256 on N GPUs).  Synthetic random-perturbation initial conditions, seed 2
(SURVEY.md §8d).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cell-updates/sec"
UNIT = "cells*steps/s"
TWO_PI = 2 * np.pi
DT = 1e-3


def workload_config(n, ft=np.float64, nx=None, ny=None, nz=None):
    """configs[1] of BASELINE.json on an nx x ny x nz grid of cubic cells (Δ = 2π/n)"""
    from helpers import Config
    nx, ny, nz = nx or n, ny or n, nz or n
    return Config((nx, ny, nz), ((0, TWO_PI * nx / n), (0, TWO_PI * ny / n), (0, TWO_PI * nz / n)), "PPP", advection=("weno", 5),
                  closure=[("scalar", 1e-3, 1e-3)], buoyancy=("tracer",), tracers=("b",), ft=ft)


# ---------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), power_w_max=max(pw))
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference CPU() path on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------------------
def cpu_run(n, steps, warmup):
    from oracle import model as M  # the one place outside tests/ and smoke() that may execute oracle/
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1: set the team size explicitly and report the size actually in effect
    cores = M.set_num_threads(avail)
    cfg = workload_config(n)
    om = cfg.oracle_model()
    om.set(**cfg.initial_conditions(2))
    for _ in range(warmup):
        om.time_step(DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        om.time_step(DT)
    dt = time.perf_counter() - t0
    return {"value": n ** 3 * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d^3 sub-domain of the workload (same physics, dt), %d RK3 steps after %d warm-up, OpenMP threads = %d; "
                      "Julia is not installed, so this is the C/numpy restatement of the reference CPU() path (oracle/)" % (n, steps, warmup, cores),
            "seconds": dt}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_size
    r = cpu_run(n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D triply-periodic NonhydrostaticModel WENO-5 BuoyancyTracer ScalarDiffusivity Float64 (configs[1]); "
                                   "CPU sample %d^3" % n, "timestepper": "RK3", "dt": DT},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=256, help="cells per side per GPU")
    ap.add_argument("--ny", type=int, default=0, help="cells in y (default: --size); configs[4] is --size 256 --ny 2048 --nz 512 on 8 GPUs")
    ap.add_argument("--nz", type=int, default=0, help="cells in z (default: --size)")
    ap.add_argument("--ref-size", type=int, default=128)
    ap.add_argument("--cpu-size", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lanes", type=int, default=3, help="device lanes (independent host-resident members) of the e2e leg; 1 = serial only")
    ap.add_argument("--overlap", action="store_true", help="distributed: compute interior tendency tiles while the x halos are in flight (OB_OPT_OVERLAP_HALO)")
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the global grid is --nx x ny x nz whatever N (default: weak, size^3 per GPU)")
    ap.add_argument("--nx", type=int, default=0, help="strong scaling: global cells in x (default: --size); the recorded runs use --strong --nx 512 --size 256")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import ocean_b200 as ob
    from ocean_b200 import _abi
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        arch = ob.Distributed(ob.B200(local))
    else:
        arch = ob.B200(local)
    ft = np.float32 if args.f32 else np.float64
    n = args.size
    ny, nz = args.ny or n, args.nz or n
    nx_global = (args.nx or n) if args.strong else n * world
    cfg = workload_config(n, ft=ft, nx=nx_global, ny=ny, nz=nz)
    nx_local = nx_global // world
    model = cfg.b200_model(arch)
    if args.overlap:
        model.set_option(_abi.OB_OPT_OVERLAP_HALO, 1)
    # synthetic random-perturbation initial conditions, generated per rank at the local size (a global array of
    # configs[4] would be 17 GB per field)
    ic = workload_config(n, ft=ft, nx=nx_local, ny=ny, nz=nz).initial_conditions(2 + rank)
    ob.set(model, **ic)
    cells_local = nx_local * ny * nz
    cells_total = cells_local * world

    def barrier():
        arch.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        _abi.call("ob_timer_start", arch.ctx)
        for _ in range(steps):
            fn()
        ms = C.c_double(0)
        _abi.call("ob_timer_stop", arch.ctx, C.byref(ms))
        arch.synchronize()
        t = ms.value
        if dist is not None:
            import torch
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
            dist.barrier()
        return t

    step = lambda: ob.time_step(model, DT)
    for _ in range(args.warmup):
        step()
    # ---- device-resident leg (value) + per-phase kernel timing (roofline) ------------------------------------
    _abi.call("ob_reset_timing", model.handle)
    _abi.call("ob_enable_timing", model.handle, 1)
    l0 = model.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop()
    launches = model.launch_count() - l0
    nph = C.c_int32(0)
    _abi.call("ob_phase_count", C.byref(nph))
    phases = {}
    for p in range(nph.value):
        t, c = C.c_double(0), C.c_int64(0)
        _abi.call("ob_phase_time_ms", model.handle, p, C.byref(t), C.byref(c))
        phases[_abi.lib().ob_phase_name(p).decode()] = {"ms_total": t.value, "calls": c.value}
    _abi.call("ob_enable_timing", model.handle, 0)
    value = cells_total * args.steps / (ms * 1e-3)
    assert not model.velocities["u"].any_nan(), "NaN in u after the timed region"

    # roofline of the dominant kernel: the fused tendency kernel; algorithmic bytes = 2(3+n) words/cell
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    tend = phases.get("tendencies", {"ms_total": 0, "calls": 0})
    wsize = np.dtype(ft).itemsize
    alg_bytes = 2 * (3 + len(cfg.tracers)) * wsize * cells_local
    roofline = None
    if tend["calls"]:
        avg_ms = tend["ms_total"] / tend["calls"]
        ach = alg_bytes / (avg_ms * 1e-3) / 1e9
        # DRAM traffic cannot be measured without a profiler: it is the figure of the committed ncu capture of this kernel
        # on this workload (profiles/tendency_traffic.json names the capture), reported only for the configuration captured
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "tendency_traffic.json")
        if os.path.exists(tpath) and (n, ny, nz) == (256, 256, 256) and not args.f32 and world == 1:
            try:
                tj = json.load(open(tpath))
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
            except Exception:
                pass
        roofline = {"kernel": "fused tendency (Gu,Gv,Gw,Gb in one launch)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "avg_launch_ms": avg_ms,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "Float64 WENO-5 is FP64-issue bound on B200, not HBM bound (DESIGN.md §roofline)"}
    # second roofline of the same kernel: FP64 issue.  FP64 thread-instructions per cell come from the committed ncu capture
    # (profiles/tendency_traffic.json: DFMA+DMUL+DADD executed / cells, tile-overlap lanes included); the peak is measured
    # live with a DFMA microbenchmark (ob_fp64_peak).
    if roofline is not None and not args.f32 and (n, ny, nz) == (256, 256, 256):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "tendency_traffic.json")))
            per_cell = tj.get("fp64_instructions_per_cell")
            pk = C.c_double(0)
            _abi.call("ob_fp64_peak", arch.ctx, C.byref(pk))
            if per_cell:
                ach = per_cell * cells_local / (roofline["avg_launch_ms"] * 1e-3)
                roofline["fp64_issue"] = {"achieved": ach / 1e12, "peak": pk.value / 1e12, "unit": "T thread-instr/s", "frac": ach / pk.value,
                                          "fp64_instructions_per_cell": per_cell, "peak_source": "measured DFMA microbenchmark (ob_fp64_peak)"}
        except Exception:
            pass
    step_bytes = 1440 * cells_local * (wsize / 8.0)  # ≈ 1.44 KB/cell/RK3 step
    step_roof = {"achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": step_bytes}

    # ---- end-to-end leg: host buffers in, host buffers out, every step ----------------------------------------
    e2e = None
    if not args.no_e2e:
        fields = list(model.prognostic_fields.values())
        pinned = []
        for f in fields:
            p = C.c_void_p()
            _abi.call("ob_malloc_host", arch.ctx, f.nbytes, C.byref(p))
            _abi.call("ob_memcpy_d2h", arch.ctx, p, f.data, f.nbytes)
            pinned.append(p)
        nbytes = sum(f.nbytes for f in fields)

        def e2e_step():
            for f, p in zip(fields, pinned):
                _abi.call("ob_memcpy_h2d", arch.ctx, f.data, p, f.nbytes)   # async on the library stream, pinned source
            ob.time_step(model, DT)
            for f, p in zip(fields, pinned):
                _abi.call("ob_memcpy_d2h", arch.ctx, p, f.data, f.nbytes)   # result back on the host

        for _ in range(3):
            e2e_step()
        ke = max(3, min(args.steps, 10))
        ms_e = timed(e2e_step, ke)
        serial = {"value": cells_total * ke / (ms_e * 1e-3), "ms_per_step": ms_e / ke, "steps": ke,
                  "what": "one stream: copy in, time_step!, copy out strictly one after the other"}
        e2e = {"value": serial["value"], "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": ms_e / ke, "steps": ke, "lanes": 1,
               "what": "per step: pinned-host -> device copy of every prognostic field (u,v,w,b parents), time_step! through the C ABI, "
                       "device -> host copy of every prognostic field"}
        if world == 1 and args.lanes > 1:
            # host-streamed ensemble (ocean_b200.HostStreamedStepper): `lanes` independent members whose states live in pinned
            # host memory; every step of every member is upload -> time_step! -> download, the lanes overlap on the device
            stepper = ob.HostStreamedStepper(lambda a: cfg.b200_model(a), lanes=args.lanes, device=local, first_arch=arch, first_model=model)
            for m in stepper.models[1:]:
                ob.set(m, **ic)
            members = [stepper.new_member() for _ in range(args.lanes)]
            for lane, mem in enumerate(members):
                stepper.download(mem, lane)

            def round_robin(nsteps):
                for s_ in range(nsteps):
                    mem = members[s_ % args.lanes]
                    stepper.step(mem, mem, DT)

            round_robin(2 * args.lanes)
            stepper.synchronize()
            kp = args.lanes * max(2, min(args.steps, 12) // args.lanes)
            _abi.call("ob_timer_start", arch.ctx)
            round_robin(kp)
            stepper.join_into(0)
            msp = C.c_double(0)
            _abi.call("ob_timer_stop", arch.ctx, C.byref(msp))
            stepper.synchronize()
            # the headline e2e number stays the ONE-model loop above; the ensemble throughput is an extra key
            e2e["ensemble"] = {"value": cells_total * kp / (msp.value * 1e-3), "ms_per_step": msp.value / kp, "steps": kp, "lanes": args.lanes,
                               "what": "%d independent ensemble members whose prognostic fields live in pinned host memory; every step of every "
                                       "member = pinned-host -> device copy, time_step! through the C ABI, device -> host copy; each member has "
                                       "its own device lane (library context = stream), so the copies of one member overlap the kernels of "
                                       "another (ocean_b200.HostStreamedStepper)" % args.lanes}
            for mem in members:
                mem.free()
            del stepper
        for p in pinned:
            _abi.call("ob_free_host", arch.ctx, p)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_run(args.cpu_size, 5, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
                "dtype": "f32" if args.f32 else "f64", "data": "synthetic",
                "config": {"workload": "3D triply-periodic NonhydrostaticModel, WENO-5, BuoyancyTracer, ScalarDiffusivity, %dx%dx%d %s "
                                       "(BASELINE.json configs[1]%s; %dx%dx%d per GPU, slab-x)" % (nx_local * world, ny, nz, "Float32" if args.f32 else "Float64",
                                                                                              " on the grid of configs[4]" if (ny, nz) != (n, n) else "", nx_local, ny, nz),
                           "timestepper": "RK3 (3 stages, 3 pressure solves per step)", "dt": DT, "halo": 3,
                           "l2": "inputs larger than L2: every kernel streams >= 4 parent arrays of %.0f MB" % (model.velocities["u"].nbytes / 1e6),
                           "parallelism": "slab-x%d" % world},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "step_roofline": step_roof,
                "phases_ms_per_step": {k: v["ms_total"] / args.steps for k, v in phases.items()},
                "e2e": e2e, "cpu_baseline": cpu}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
