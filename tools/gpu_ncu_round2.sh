#!/bin/bash
# profiler visit (one GPU), sized to finish in ~4 minutes and to bring back < 20 MB: `ncu --set full` of the staged-ring kernel
# (config 2) and of the LES kernels at 256^3 (config-3 physics: tendency MN / TT, amd, hydrostatic pressure, DCT solver kernels;
# config-4 physics: thomas), converted to raw CSV on the box (the .ncu-rep files stay there)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
cap() {  # cap <tag> <kernel regex> <count> <command...>
  local tag=$1 rx=$2 cnt=$3; shift 3
  timeout 240 ncu --set full --clock-control none -k "regex:$rx" --launch-skip 6 --launch-count $cnt -f -o /tmp/$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
  echo "$tag rc=$? $(wc -c < gpurun_out/${tag}_raw.csv) bytes"
}
cap r2_stage_final 'tendency_stage' 1 env OB_MODES=8 OB_FT=f64 python tools/bench_tendency.py 256 4
cap r2_les_tend 'tendency_stage|tendency_march' 3 python tools/les_step.py 256 regular
cap r2_les_amd 'amd_kernel' 1 python tools/les_step.py 256 regular
cap r2_les_misc 'hydrostatic_pressure|dct_z_real|update_vec|source_pair|correct_pair|halo_kernel|flux_bc' 10 python tools/les_step.py 256 regular
cap r2_les_thomas 'thomas_kernel' 1 python tools/les_step.py 256 stretched
ls -la gpurun_out/*_raw.csv | tail -6
