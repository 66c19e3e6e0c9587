"""World-size-2 gloo check of the host-side distributed logic (no GPU): partitioning, neighbours, scalar reductions,
gathers, per-rank grid windows.  Run under torchrun by tests/test_distributed_host.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import ocean_b200 as ob  # noqa: E402
from ocean_b200.grids import RectilinearGrid  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()


class FakeArch:  # the host logic only needs rank / world (no device in this container)
    def __init__(self):
        self.rank, self.world, self.ctx = rank, world, None


arch = FakeArch()
nx, x0 = ob.partition_x(32, world, rank)
assert (nx, x0) == (16, 16 * rank)
assert ob.neighbors(rank, world) == ((rank - 1) % world, (rank + 1) % world)
assert ob.all_reduce_scalar(arch, 3.0 + rank, "min") == 3.0
assert ob.all_reduce_scalar(arch, 3.0 + rank, "max") == 3.0 + world - 1
g = ob.gather_x(arch, np.full((2, 3, 4), float(rank)))
assert g.shape == (2, 3, 4 * world) and all(np.all(g[:, :, 4 * r:4 * r + 4] == r) for r in range(world))
# per-rank grid = window of the global grid with the SAME spacing on every rank
grid = RectilinearGrid(arch, size=(32, 8, 8), x=(0, 2 * np.pi), y=(0, 1), z=(0, 1), topology=(ob.Periodic, ob.Periodic, ob.Periodic))
solo = FakeArch(); solo.world, solo.rank = 1, 0
full = RectilinearGrid(solo, size=(32, 8, 8), x=(0, 2 * np.pi), y=(0, 1), z=(0, 1), topology=(ob.Periodic, ob.Periodic, ob.Periodic))
assert grid.N == (16, 8, 8) and grid.N_global == (32, 8, 8)
assert grid.dF[0] == full.dF[0]
assert np.array_equal(grid.nodes(0, "c"), full.nodes(0, "c")[16 * rank:16 * rank + 16])
assert np.array_equal(grid.nodes(0, "f"), full.nodes(0, "f")[16 * rank:16 * rank + 16])
assert abs(float(grid.L[0]) * world - float(full.L[0])) < 1e-15
try:
    ob.partition_x(33, world, rank)
    raise SystemExit("partition_x accepted an indivisible size")
except ValueError:
    pass
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
