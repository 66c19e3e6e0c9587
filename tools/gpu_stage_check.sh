#!/bin/bash
# GPU check of the staged-ring tendency kernel: sanitizer, bit-identity, oracle parity, timing (config 2 and LES)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "staged_ring and (stage_ppp or stage_les or wide_stretched)" > gpurun_out/san_mem.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/san_mem.log
echo "== bit identity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "staged_ring" > gpurun_out/t_stage.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_stage.log
echo "== oracle parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tendencies_match and stage" > gpurun_out/t_stage2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_stage2.log
echo "== timing"; timeout 600 env OB_MODES=2,8 python tools/bench_tendency.py 256 20 > gpurun_out/bt.log 2>&1; echo "rc=$?"; cat gpurun_out/bt.log
timeout 600 env OB_CASE=les OB_MODES=2,8 python tools/bench_tendency.py 256 10 2>&1 | tail -4
