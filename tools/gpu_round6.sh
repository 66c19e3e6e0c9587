#!/bin/bash
# 4-GPU visit: distributed parity at world 4, weak scaling N = 4, strong scaling (512 x 256 x 256 global) at N = 1, 2, 4
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2h}
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
echo "== dist_check w4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 tests/dist_check.py > gpurun_out/${TAG}_dist_check_w4.log 2>&1; echo "rc=$?" | tee -a gpurun_out/${TAG}_dist_check_w4.log; grep -v "^W1\|warn\|Warn\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_dist_check_w4.log | tail -13
echo "== weak n4"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 4 --no-cpu --no-e2e > gpurun_out/bench_${TAG}_n4.json 2> gpurun_out/bench_${TAG}_n4.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n4.json
echo "== strong n1"; timeout 600 python bench.py --gpus 1 --strong --nx 512 --size 256 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_strong_n1.json 2> gpurun_out/bench_${TAG}_strong_n1.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_strong_n1.json
for N in 2 4; do
echo "== strong n$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29650+N)) bench.py --gpus $N --strong --nx 512 --size 256 --no-e2e --no-cpu > gpurun_out/bench_${TAG}_strong_n$N.json 2> gpurun_out/bench_${TAG}_strong_n$N.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_strong_n$N.json
done
echo "== amd variants"; for mb in 4 5 6; do OB_AMD_MINB=$mb python tools/bench_closure.py 256 10 2>&1 | head -1; done
