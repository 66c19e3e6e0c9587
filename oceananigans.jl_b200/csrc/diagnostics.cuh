// diagnostics.cuh -- device reductions used by the Simulation driver around the hot path:
//   cell_advection_timescale  src/Advection/cell_advection_timescale.jl:14-35  (TimeStepWizard)
// Included at the end of ocean_b200.cu (needs ModelT).
#pragma once

namespace ob {

template <typename T>
struct CflP {
    GridD<T> g;
    Fld<T> u, v, w;
};

// min over the interior of 1 / (|u|/Δxᶠ + |v|/Δyᶠ + |w|/Δzᶠ); one partial minimum per block
template <typename T>
__global__ void __launch_bounds__(256) advection_timescale_kernel(const __grid_constant__ CflP<T> P, double *partial) {
    const GridD<T> &g = P.g;
    const long n = (long)g.N[0] * g.N[1] * g.N[2];
    double best = INFINITY;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const int i = 1 + (int)(t % g.N[0]), j = 1 + (int)((t / g.N[0]) % g.N[1]), k = 1 + (int)(t / ((long)g.N[0] * g.N[1]));
        const T ix = g.topo[0] == FLAT ? T(0) : fabs(P.u.ld(i, j, k)) * g.rdx;
        const T iy = g.topo[1] == FLAT ? T(0) : fabs(P.v.ld(i, j, k)) * g.rdy;
        const T iz = g.topo[2] == FLAT ? T(0) : fabs(P.w.ld(i, j, k)) * g.rdzF(k);
        const T inv = ix + iy + iz;
        best = fmin(best, (double)(1 / inv));
    }
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; q++) best = fmin(best, s[q]);
        partial[blockIdx.x] = best;
    }
}
__global__ void min_reduce_kernel(const double *partial, int n, double *out) {
    double best = INFINITY;
    for (int t = threadIdx.x; t < n; t += blockDim.x) best = fmin(best, partial[t]);
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; q++) best = fmin(best, s[q]);
        *out = best;
    }
}

}  // namespace ob

template <typename T>
int32_t ModelT<T>::advection_timescale(double *tau) {
    OB_TRY(need(OB_FIELD_U)); OB_TRY(need(OB_FIELD_V)); OB_TRY(need(OB_FIELD_W));
    CflP<T> P;
    memset(&P, 0, sizeof(P));
    P.g = g; P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W);
    const int nblocks = ctx->sm_count * 4;
    if (!d_partial) CUDA_TRY(cudaMalloc(&d_partial, sizeof(double) * (nblocks + 1)));
    advection_timescale_kernel<T><<<nblocks, 256, 0, ctx->stream>>>(P, d_partial);
    min_reduce_kernel<<<1, 256, 0, ctx->stream>>>(d_partial, nblocks, d_partial + nblocks);
    launches += 2;
    double h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d_partial + nblocks, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *tau = h;
    return OB_OK;
}
