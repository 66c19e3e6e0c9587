#!/bin/bash
# one ncu --set full capture of the staged-ring tendency kernel at 256^3 (F64), source counters included
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
MODE=${1:-8}; TAG=${2:-stage}
timeout 180 env OB_MODES=$MODE OB_FT=f64 ncu --set full --import-source on --clock-control none -k regex:tendency_ --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/$TAG python tools/bench_tendency.py 256 2 > gpurun_out/ncu_$TAG.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/$TAG.ncu-rep
