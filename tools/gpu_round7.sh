#!/bin/bash
# final all-GPU visit: distributed parity, weak scaling (default + chunk-count variants), strong scaling
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TAG=${1:-r2k}
echo "== dist_check w$NG"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29661 tests/dist_check.py > gpurun_out/${TAG}_dist_check_w$NG.log 2>&1; echo "rc=$?" | tee -a gpurun_out/${TAG}_dist_check_w$NG.log; grep -v "^W1\|warn\|Warn\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_dist_check_w$NG.log | tail -13
echo "== bench n$NG (default)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus $NG --no-cpu > gpurun_out/bench_${TAG}_n$NG.json 2> gpurun_out/bench_${TAG}_n$NG.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n$NG.json
for CH in 2 8; do
echo "== bench n$NG, OB_DIST_CHUNKS=$CH"; OB_DIST_CHUNKS=$CH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29670+CH)) bench.py --gpus $NG --no-e2e --no-cpu > gpurun_out/bench_${TAG}_n${NG}_chunks$CH.json 2> gpurun_out/bench_${TAG}_n${NG}_chunks$CH.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n${NG}_chunks$CH.json
done
echo "== strong scaling 512x256x256 global on $NG"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29664 bench.py --gpus $NG --strong --nx 512 --size 256 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_strong_n$NG.json 2> gpurun_out/bench_${TAG}_strong_n$NG.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_strong_n$NG.json
