#!/bin/bash
# all-GPU visit: distributed parity at world = all visible GPUs, weak-scaling bench with / without the pipelined solve, strong scaling line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TAG=${1:-r2f}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== dist_check w$NG"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29631 tests/dist_check.py > gpurun_out/${TAG}_dist_check_w$NG.log 2>&1; echo "rc=$?" | tee -a gpurun_out/${TAG}_dist_check_w$NG.log; grep -v "^W1\|warn\|Warn\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_dist_check_w$NG.log | tail -13
echo "== bench n$NG (default)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $NG --no-cpu > gpurun_out/bench_${TAG}_n$NG.json 2> gpurun_out/bench_${TAG}_n$NG.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n$NG.json
echo "== bench n$NG, OB_DIST_PIPELINE=0"; OB_DIST_PIPELINE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus $NG --no-e2e --no-cpu > gpurun_out/bench_${TAG}_n${NG}_nopipe.json 2> gpurun_out/bench_${TAG}_n${NG}_nopipe.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n${NG}_nopipe.json
echo "== strong scaling 512x256x256 global on $NG"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus $NG --strong --nx 512 --size 256 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_strong_n$NG.json 2> gpurun_out/bench_${TAG}_strong_n$NG.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_strong_n$NG.json
