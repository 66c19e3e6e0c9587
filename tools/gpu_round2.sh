#!/bin/bash
# 2-GPU visit: single-GPU test suite on GPU 0, distributed parity at world 2, bench at N = 1 and N = 2
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/gpu_all.log
echo "== dist_check w2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_check.py > gpurun_out/r2c_dist_check_w2.log 2>&1; echo "rc=$?"; grep -v "^W1\|warn\|Warn" gpurun_out/r2c_dist_check_w2.log | tail -14
echo "== bench n1"; timeout 600 python bench.py > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_r2c_n1.json
echo "== bench n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_r2c_n2.json
