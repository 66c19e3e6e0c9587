#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
OB_CASE=les OB_MODES=2,8 OB_FT=f64 python tools/bench_tendency.py 256 10 2>&1 | tail -3
timeout 300 env OB_CASE=les OB_MODES=8 OB_FT=f64 ncu --set full --import-source on --clock-control none -k regex:tendency_stage --launch-skip 3 --launch-count 1 -f -o gpurun_out/les_stage python tools/bench_tendency.py 256 2 > gpurun_out/ncu_les.log 2>&1; tail -2 gpurun_out/ncu_les.log
timeout 300 env OB_CASE=les OB_MODES=8 OB_FT=f64 ncu --set full --clock-control none -k regex:amd_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/les_amd python tools/bench_tendency.py 256 2 > gpurun_out/ncu_amd.log 2>&1; tail -2 gpurun_out/ncu_amd.log
