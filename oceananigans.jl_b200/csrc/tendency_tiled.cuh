// tendency_tiled.cuh -- shared-memory tiled fast path of the fused tendency kernel (placeholder dispatcher:
// returns done = false so the generic kernel runs).
#pragma once
#include "tendency.cuh"

namespace ob {
template <typename T, class S>
static cudaError_t try_tiled_tendency(const TendP<T> &P, int fast, cudaStream_t st, int sm_count, int *nlaunch, bool &done) {
    (void)P; (void)fast; (void)st; (void)sm_count; (void)nlaunch;
    done = false;
    return cudaSuccess;
}
}  // namespace ob
