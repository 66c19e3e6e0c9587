#!/bin/bash
# one GPU visit: staged-kernel checks, full GPU test suite, bench line, per-config table
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
bash tools/gpu_stage_check.sh
echo "== gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/gpu_all.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2b.json
echo "== configs"; timeout 900 python tools/bench_configs.py > gpurun_out/r2b_config_table.jsonl 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2b_config_table.jsonl
