// tend_launch.h -- host-side dispatch to the fused tendency kernels.  Every (float type, scheme kind, buffer) is
// compiled in its own translation unit (tend_inst.cu with -DOB_TI_*), so the library builds in parallel.
#pragma once
#include "tendency.cuh"

namespace ob {
// Signature of one instantiation's launcher; returns cudaErrorNotSupported when the combination is not built.
#define OB_TEND_DECL(T, TN, KIND, NB) \
    cudaError_t launch_tend_##TN##_k##KIND##_n##NB(const TendP<T> &P, int fast, int mode, cudaStream_t st, int sm_count, int *nlaunch, int tx_lo, int tx_hi, int invert);
#define OB_TEND_ALL(X) \
    X(double, f64, 0, 0) X(float, f32, 0, 0) \
    X(double, f64, 1, 1) X(double, f64, 1, 2) X(double, f64, 1, 3) X(double, f64, 1, 4) X(double, f64, 1, 5) X(double, f64, 1, 6) \
    X(float, f32, 1, 1) X(float, f32, 1, 2) X(float, f32, 1, 3) X(float, f32, 1, 4) X(float, f32, 1, 5) X(float, f32, 1, 6) \
    X(double, f64, 2, 2) X(double, f64, 2, 3) X(double, f64, 2, 4) X(double, f64, 2, 5) X(double, f64, 2, 6) \
    X(float, f32, 2, 2) X(float, f32, 2, 3) X(float, f32, 2, 4) X(float, f32, 2, 5) X(float, f32, 2, 6)
OB_TEND_ALL(OB_TEND_DECL)

inline cudaError_t launch_tendency(const TendP<double> &P, int kind, int nb, int fast, int mode, cudaStream_t st, int sm, int *nl, int tx_lo = 0, int tx_hi = -1, int invert = 0) {
#define OB_CASE64(T, TN, KIND, NB) if (sizeof(T) == 8 && kind == KIND && nb == NB) return launch_tend_f64_k##KIND##_n##NB(P, fast, mode, st, sm, nl, tx_lo, tx_hi, invert);
    OB_CASE64(double, f64, 0, 0) OB_CASE64(double, f64, 1, 1) OB_CASE64(double, f64, 1, 2) OB_CASE64(double, f64, 1, 3)
    OB_CASE64(double, f64, 1, 4) OB_CASE64(double, f64, 1, 5) OB_CASE64(double, f64, 1, 6)
    OB_CASE64(double, f64, 2, 2) OB_CASE64(double, f64, 2, 3) OB_CASE64(double, f64, 2, 4) OB_CASE64(double, f64, 2, 5) OB_CASE64(double, f64, 2, 6)
    return cudaErrorNotSupported;
}
inline cudaError_t launch_tendency(const TendP<float> &P, int kind, int nb, int fast, int mode, cudaStream_t st, int sm, int *nl, int tx_lo = 0, int tx_hi = -1, int invert = 0) {
#define OB_CASE32(T, TN, KIND, NB) if (sizeof(T) == 4 && kind == KIND && nb == NB) return launch_tend_f32_k##KIND##_n##NB(P, fast, mode, st, sm, nl, tx_lo, tx_hi, invert);
    OB_CASE32(float, f32, 0, 0) OB_CASE32(float, f32, 1, 1) OB_CASE32(float, f32, 1, 2) OB_CASE32(float, f32, 1, 3)
    OB_CASE32(float, f32, 1, 4) OB_CASE32(float, f32, 1, 5) OB_CASE32(float, f32, 1, 6)
    OB_CASE32(float, f32, 2, 2) OB_CASE32(float, f32, 2, 3) OB_CASE32(float, f32, 2, 4) OB_CASE32(float, f32, 2, 5) OB_CASE32(float, f32, 2, 6)
    return cudaErrorNotSupported;
}
}  // namespace ob
