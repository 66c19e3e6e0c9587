"""Host-streamed time stepping: model states that live in (pinned) host memory are advanced through the GPU.

The drop-in boundary of this library is a C ABI over *device* fields; a host program whose states live in host memory
(an ensemble larger than HBM, or a host-side coupler that owns the prognostic arrays between steps) pays one
host->device copy of every prognostic parent array before `time_step!` and one device->host copy after it.  Issued on
one stream those three phases serialise (copy in, ~30 kernels, copy out); `HostStreamedStepper` keeps `lanes` device
replicas of the model, each on its own library context (= its own CUDA stream, cuFFT plans and workspace), and hands
the members to the lanes round-robin, so that the upload of member n+1 and the download of member n-1 overlap the
kernels of member n (the two PCIe directions and the SMs are three independent resources).

There is no reference counterpart: Oceananigans' CPU() path has no host/device boundary and its GPU path keeps the
state resident (as `ob.time_step` does).  bench.py's `e2e` leg is measured through this class.
"""
import ctypes as C

from . import _abi
from .grids import B200
from .models import time_step


class HostMember:
    """One model state in pinned host memory: a buffer per prognostic parent array (u, v, w, tracers...)."""

    def __init__(self, arch, fields):
        self.arch = arch
        self.nbytes = [f.nbytes for f in fields]
        self.lane = None          # the device lane this member is bound to (set by its first step)
        self.clock = None         # (time, iteration, last_dt, last_stage_dt) of this member, carried between steps
        self.ptrs = []
        for f in fields:
            p = C.c_void_p()
            _abi.call("ob_malloc_host", arch.ctx, f.nbytes, C.byref(p))
            self.ptrs.append(p)

    def free(self):
        for p in self.ptrs:
            _abi.call("ob_free_host", self.arch.ctx, p)
        self.ptrs = []

    def total_bytes(self):
        return sum(self.nbytes)


class HostStreamedStepper:
    """`HostStreamedStepper(make_model, lanes=3, device=0)`: `make_model(arch)` builds the model on a fresh `B200`
    context; it is called once per lane.  `step(member_in, member_out, dt)` enqueues upload -> time_step! -> download
    on the next lane and returns immediately; `synchronize()` waits for every lane."""

    def __init__(self, make_model, lanes=3, device=0, first_arch=None, first_model=None):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.archs, self.models = [], []
        for n in range(lanes):
            if n == 0 and first_model is not None:
                arch, model = first_arch, first_model
            else:
                arch = B200(device)
                model = make_model(arch)
            self.archs.append(arch)
            self.models.append(model)
        self._next = 0
        self._holder = [None] * lanes   # the member whose state (and tendencies) each lane holds

    @property
    def lanes(self):
        return len(self.models)

    def new_member(self, like_lane=0):
        return HostMember(self.archs[0], list(self.models[like_lane].prognostic_fields.values()))

    def download(self, member, lane=0):
        """blocking copy of lane's current device state into `member` (initialisation of host states)"""
        arch = self.archs[lane]
        for f, p in zip(self.models[lane].prognostic_fields.values(), member.ptrs):
            _abi.call("ob_memcpy_d2h", arch.ctx, p, f.data, f.nbytes)

    def step(self, member_in, member_out, dt):
        """One step of `member_in`; the result lands in `member_out` (usually the same object).  A member is bound to the
        lane of its first step, so its uploads and downloads are ordered on one stream.  When a lane receives a member
        other than the one it held last (more members than lanes), the tendencies left on the lane belong to another
        state: `update_state!` is re-run after the upload and the member's own clock is installed.  Adams-Bashforth
        models are refused in that case: G⁻ is part of their state and is not carried in a HostMember."""
        if member_in.lane is None:
            member_in.lane = self._next
            self._next = (self._next + 1) % self.lanes
        lane = member_in.lane
        if member_out.lane is None:
            member_out.lane = lane
        if member_out.lane != lane:
            raise ValueError("member_out is bound to lane %d, member_in to lane %d: a member's copies must stay on one stream" % (member_out.lane, lane))
        arch, model = self.archs[lane], self.models[lane]
        fields = list(model.prognostic_fields.values())
        for f, p in zip(fields, member_in.ptrs):
            _abi.call("ob_memcpy_h2d", arch.ctx, f.data, p, f.nbytes)        # async, pinned source
        clk = model.clock
        if self._holder[lane] is not member_in:
            if member_in.clock is not None:
                clk.time, clk.iteration, clk.last_dt, clk.last_stage_dt = member_in.clock
            if clk.iteration != 0:   # (at iteration 0 the step itself starts with update_state!)
                if model.timestepper != "RungeKutta3":
                    raise NotImplementedError("sharing a lane between members needs G⁻ in the host state (QuasiAdamsBashforth2)")
                model.update_state()   # Gⁿ of THIS member, not of the member the lane held before
        time_step(model, dt)                                                  # ONE C-ABI call, async on the lane's stream
        member_out.clock = (clk.time, clk.iteration, clk.last_dt, clk.last_stage_dt)
        for f, p in zip(fields, member_out.ptrs):
            _abi.call("ob_memcpy_d2h_async", arch.ctx, p, f.data, f.nbytes)  # stream-ordered, pinned destination
        self._holder[lane] = member_out
        return lane

    def join_into(self, lane=0):
        """make lane `lane`'s stream wait (on the device) for all other lanes: an event recorded on it afterwards
        follows every member submitted so far"""
        for n, a in enumerate(self.archs):
            if n != lane:
                _abi.call("ob_stream_wait", self.archs[lane].ctx, a.ctx)

    def synchronize(self):
        for a in self.archs:
            a.synchronize()
