"""Single-node multi-GPU: `Distributed(B200(local_rank))` -- slab-x domain decomposition, one process per GPU.

Reference: src/DistributedComputations/distributed_architectures.jl:164-305 (`Distributed(child; partition)`,
`Partition(R)`, rank <-> index maps :352-361, periodic rank connectivity :407-), ext/OceananigansNCCLExt/
nccl_communicator.jl:25-63 (bootstrap: rank 0 creates the NCCL unique id, the host broadcasts it -- MPI there,
`torch.distributed` here -- and every rank calls ncclCommInitRank).  torch.distributed is plumbing only: after the
bootstrap every halo exchange and FFT transpose runs inside libocean_b200.so on NCCL over NVLink.
"""
import ctypes as C

import numpy as np

from . import _abi


def partition_x(Nx, world, rank):
    """Partition(R) in x: equal slabs (distributed_architectures.jl:240-305); returns (nx, first global index, 0-based)"""
    if Nx % world:
        raise ValueError("Nx = %d must be divisible by the number of ranks %d (equal slab-x partition)" % (Nx, world))
    nx = Nx // world
    return nx, rank * nx


def neighbors(rank, world):
    """periodic west / east neighbours (NeighboringRanks, distributed_architectures.jl:407-)"""
    return (rank - 1) % world, (rank + 1) % world


def Distributed(arch, partition=None):
    """Attach the slab-x communicator to a B200 architecture.  `torch.distributed` must be initialised (any backend)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("Distributed(B200()) needs torch.distributed.init_process_group (one process per GPU)")
    rank, world = dist.get_rank(), dist.get_world_size()
    if partition is not None and int(partition) != world:
        raise ValueError("Partition(%s) does not match the world size %d" % (partition, world))
    ident = [None]
    if rank == 0:
        buf = (C.c_char * 128)()
        _abi.call("ob_dist_unique_id", C.cast(buf, C.c_void_p))
        ident[0] = bytes(buf)
    dist.broadcast_object_list(ident, src=0)
    buf = (C.c_char * 128).from_buffer_copy(ident[0])
    _abi.call("ob_dist_init", arch.ctx, rank, world, C.cast(buf, C.c_void_p))
    arch.rank, arch.world = rank, world
    arch.distributed = True
    return arch


def all_reduce_scalar(arch, value, op):
    """host-side scalar reduction over ranks (time_step_wizard.jl `all_reduce(min, ...)`, run.jl:191-197 NaN flag)"""
    if getattr(arch, "world", 1) == 1:
        return value
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op={"min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
    return float(t.item())


def gather_x(arch, local):
    """concatenate per-rank interior arrays (numpy, shape (nz, ny, nx_local)) along x on every rank"""
    if getattr(arch, "world", 1) == 1:
        return local
    import torch.distributed as dist
    parts = [None] * arch.world
    dist.all_gather_object(parts, np.ascontiguousarray(local))
    return np.concatenate(parts, axis=2)
