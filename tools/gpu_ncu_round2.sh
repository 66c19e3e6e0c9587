#!/bin/bash
# profiler visit (one GPU): launch list of the bench command, full capture of the staged-ring kernel, full captures of the LES
# kernels (config-3 physics at 256^3: amd, tendency MN / TT, dct, thomas via the stretched case) -- summarised under profiles/
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_launches_bench.log 2>&1; echo "rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launch_list.txt 2>&1; head -16 gpurun_out/r2_launch_list.txt
echo "== staged-ring kernel, config 2"
timeout 300 env OB_MODES=8 OB_FT=f64 ncu --set full --import-source on --clock-control none -k regex:tendency_stage --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_stage_final python tools/bench_tendency.py 256 2 > gpurun_out/ncu_r2_stage_final.log 2>&1; tail -2 gpurun_out/ncu_r2_stage_final.log
echo "== one LES step (config-3 physics, 256^3): every kernel"
timeout 900 ncu --set full --clock-control none --launch-skip 130 --launch-count 110 -f -o gpurun_out/r2_les_step python tools/les_step.py 256 regular > gpurun_out/ncu_r2_les_step.log 2>&1; tail -2 gpurun_out/ncu_r2_les_step.log
echo "== one LES step on stretched z (config-4 physics, 256x256x128): every kernel"
timeout 900 ncu --set full --clock-control none --launch-skip 130 --launch-count 110 -f -o gpurun_out/r2_les_stretched_step python tools/les_step.py 256 stretched > gpurun_out/ncu_r2_les_stretched.log 2>&1; tail -2 gpurun_out/ncu_r2_les_stretched.log
ls -la gpurun_out/*.ncu-rep | tail -5
