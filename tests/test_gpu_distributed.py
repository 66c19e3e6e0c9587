"""Multi-GPU (slab-x) parity: launches tests/dist_check.py under torchrun with 2 ranks when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_slab_x_matches_single_gpu():
    import ctypes as C
    import ocean_b200 as ob
    n = C.c_int32(0)
    ob.lib().ob_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
