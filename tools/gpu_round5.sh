#!/bin/bash
# one-GPU visit: full GPU test suite, closure-kernel timing, bench line, per-config table
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
TAG=${1:-r2g}
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/gpu_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/gpu_all.log
echo "== closure"; timeout 300 python tools/bench_closure.py 256 10 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "rc=$?"; python tools/bench_brief.py gpurun_out/bench_${TAG}_n1.json
echo "== configs"; timeout 900 python tools/bench_configs.py > gpurun_out/${TAG}_config_table.jsonl 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/${TAG}_config_table.jsonl
