// tend_inst.cu -- one instantiation of the fused tendency kernels; compiled once per (float type, scheme kind,
// buffer) with -DOB_TI_T=double -DOB_TI_TN=f64 -DOB_TI_KIND=2 -DOB_TI_NB=3 (see oceananigans.jl_b200/build.py).
#include "tables.cuh"
#include "tendency.cuh"
#include "tendency_tiled.cuh"

namespace ob {

#define OB_CAT_(a, b, c, d) launch_tend_##a##_k##b##_n##c
#define OB_CAT(a, b, c) OB_CAT_(a, b, c, 0)

cudaError_t OB_CAT(OB_TI_TN, OB_TI_KIND, OB_TI_NB)(const TendP<OB_TI_T> &P, int fast, int mode, cudaStream_t st, int sm_count, int *nlaunch, int tx_lo, int tx_hi, int invert) {
    using T = OB_TI_T;
    using S = Scheme<OB_TI_KIND, OB_TI_NB>;
    cudaError_t e = upload_tables();
    if (e != cudaSuccess) return e;
    bool done = false;
    e = try_tiled_tendency<T, S>(P, fast, mode, st, sm_count, nlaunch, done, tx_lo, tx_hi, invert);
    if (e != cudaSuccess) return e;
    if (!done) {
        if (tx_lo != 0 || tx_hi >= 0) return cudaErrorInvalidValue;   // the one-thread-per-cell kernel is never split
        const int Nx = P.g.N[0];
        const int bs = Nx >= 128 ? 128 : Nx >= 64 ? 64 : 32;
        dim3 grid((unsigned)((Nx + bs - 1) / bs) * (unsigned)P.g.N[1] * (unsigned)P.g.N[2], 3 + P.ntr);
        if (OB_TI_KIND == ADV_WENO && fast) tendency_generic_kernel<T, S, true><<<grid, bs, 0, st>>>(P, 1, Nx);
        else tendency_generic_kernel<T, S, false><<<grid, bs, 0, st>>>(P, 1, Nx);
        *nlaunch += 1;
    }
    return cudaGetLastError();
}

}  // namespace ob
