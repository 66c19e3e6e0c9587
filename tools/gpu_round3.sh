#!/bin/bash
# multi-GPU visit: distributed parity at world = all visible GPUs, bench with and without the pipelined solve
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TAG=${1:-r2d}
echo "== dist_check w$NG"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29621 tests/dist_check.py > gpurun_out/${TAG}_dist_check_w$NG.log 2>&1; echo "rc=$?" | tee -a gpurun_out/${TAG}_dist_check_w$NG.log; grep -v "^W1\|warn\|Warn\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_dist_check_w$NG.log | tail -14
echo "== bench n$NG"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $NG > gpurun_out/bench_${TAG}_n$NG.json 2> gpurun_out/bench_${TAG}_n$NG.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n$NG.json
echo "== bench n$NG, solve not pipelined"; OB_DIST_NO_PIPELINE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus $NG --no-e2e > gpurun_out/bench_${TAG}_n${NG}_nopipe.json 2> gpurun_out/bench_${TAG}_n${NG}_nopipe.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n${NG}_nopipe.json
