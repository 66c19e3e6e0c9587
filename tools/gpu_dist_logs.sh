#!/bin/bash
# multi-GPU parity logs: tests/dist_check.py at world = 2 .. N (N = visible GPUs), kept under profiles/
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for W in 2 4 8; do
  [ $W -le $NG ] || continue
  echo "== dist_check world=$W"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29600+W)) tests/dist_check.py > gpurun_out/r2_dist_check_w$W.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/r2_dist_check_w$W.log
  grep -v "^W1\|warn\|Warn" gpurun_out/r2_dist_check_w$W.log | tail -15
done
