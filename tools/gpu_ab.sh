#!/bin/bash
# A/B on one box: tendency kernel of a variant build of the library (OCEAN_B200_LIB) against the current one, interleaved
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
ALT=${1:-oceananigans.jl_b200/ab_w12.so}
for rep in 1 2; do
  [ -f $ALT ] && { echo "-- variant $ALT"; OCEAN_B200_LIB=$PWD/$ALT OB_MODES=8 python tools/bench_tendency.py 256 30 2>&1 | tail -2; }
  echo "-- current"; OB_MODES=8 python tools/bench_tendency.py 256 30 2>&1 | tail -2
done
[ -f $ALT ] && { echo "-- variant, LES"; OCEAN_B200_LIB=$PWD/$ALT OB_CASE=les OB_MODES=8 OB_FT=f64 python tools/bench_tendency.py 256 10 2>&1 | tail -1; }
echo "-- current, LES"; OB_CASE=les OB_MODES=8 OB_FT=f64 python tools/bench_tendency.py 256 10 2>&1 | tail -1
