"""Multi-GPU (slab-x) parity: launches tests/dist_check.py under torchrun with 2, 4 and 8 ranks (as many as GPUs are visible):
distributed model vs the single-GPU model of the same global problem -- halos and tendencies bit-identical, three time
steps <= 1e-11, distributed Poisson solvers <= 1e-11 (test/test_distributed_models.jl:626-643,
test_distributed_poisson_solvers.jl:31-135).  The logs of the builder's own 2/4/8-rank runs are under profiles/."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_x_matches_single_gpu(world):
    import ctypes as C
    import ocean_b200 as ob
    n = C.c_int32(0)
    ob.lib().ob_device_count(C.byref(n))
    if n.value < world:
        pytest.skip("needs >= %d GPUs (run with gpurun --gpus %d)" % (world, world))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29517 + world), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
