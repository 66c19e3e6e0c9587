#!/bin/bash
# A/B on one box: tendency kernel of an older build of the library (OCEAN_B200_LIB) against the current one, interleaved
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OLD=oceananigans.jl_b200/ab_c06c61f.so
for rep in 1 2; do
  [ -f $OLD ] && { echo "-- old"; OCEAN_B200_LIB=$PWD/$OLD OB_MODES=8 OB_FT=f64 python tools/bench_tendency.py 256 30 2>&1 | tail -1; }
  echo "-- new"; OB_MODES=8 OB_FT=f64 python tools/bench_tendency.py 256 30 2>&1 | tail -1
done
