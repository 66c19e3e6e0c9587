// stage_launch.h -- host-side dispatch to the staged-ring tendency kernels (tendency_stage.cuh).  Every (float type, buffer,
// mode, closure count, eddy-viscosity kind) is compiled in its own translation unit (stage_inst.cu with -DOB_SI_*), so the
// library builds in parallel; this header only declares the launchers.
#pragma once
#include "tendency.cuh"

namespace ob {
enum { STAGE_MODE_MT = 0, STAGE_MODE_MN = 1, STAGE_MODE_TT = 2 };
// X(T, TN, N, MODE, NCL, KL): MT (all closures ScalarDiffusivity) with 0 or 1 closures; MN / TT (eddy-viscosity closure KL = 2
// Smagorinsky, 3 AMD in first position) with 1 or 2 closures
#define OB_STAGE_VARIANTS(X, T, TN) \
    X(T, TN, 3, 0, 0, 0) X(T, TN, 3, 0, 1, 0) \
    X(T, TN, 3, 1, 1, 2) X(T, TN, 3, 1, 1, 3) X(T, TN, 3, 1, 2, 2) X(T, TN, 3, 1, 2, 3) \
    X(T, TN, 3, 2, 1, 2) X(T, TN, 3, 2, 1, 3) X(T, TN, 3, 2, 2, 2) X(T, TN, 3, 2, 2, 3)
#define OB_STAGE_DECL(T, TN, N, MODE, NCL, KL) \
    cudaError_t launch_stage_##TN##_n##N##_m##MODE##_c##NCL##_k##KL(const TendP<T> &P, cudaStream_t st, int sm_count, int *nlaunch);
OB_STAGE_VARIANTS(OB_STAGE_DECL, double, f64)
OB_STAGE_VARIANTS(OB_STAGE_DECL, float, f32)

inline cudaError_t launch_stage_variant(const TendP<double> &P, int n, int mode, int ncl, int kl, cudaStream_t st, int sm, int *nl) {
#define OB_STAGE_CASE64(T, TN, N, MODE, NCL, KL) if (n == N && mode == MODE && ncl == NCL && kl == KL) return launch_stage_f64_n##N##_m##MODE##_c##NCL##_k##KL(P, st, sm, nl);
    OB_STAGE_VARIANTS(OB_STAGE_CASE64, double, f64)
    return cudaErrorNotSupported;
}
inline cudaError_t launch_stage_variant(const TendP<float> &P, int n, int mode, int ncl, int kl, cudaStream_t st, int sm, int *nl) {
#define OB_STAGE_CASE32(T, TN, N, MODE, NCL, KL) if (n == N && mode == MODE && ncl == NCL && kl == KL) return launch_stage_f32_n##N##_m##MODE##_c##NCL##_k##KL(P, st, sm, nl);
    OB_STAGE_VARIANTS(OB_STAGE_CASE32, float, f32)
    return cudaErrorNotSupported;
}
}  // namespace ob
