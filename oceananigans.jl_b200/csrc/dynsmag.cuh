// dynsmag.cuh -- DynamicSmagorinsky with a directionally averaged coefficient.
//
// Reference: src/TurbulenceClosures/turbulence_closure_implementations/Smagorinskys/
//   dynamic_coefficient.jl:245-351       square_smagorinsky_coefficient, _compute_Σ!, _compute_Σ̄!, _compute_LM_MM!, LM_and_MM,
//                                        compute_coefficient_fields!(::DirectionallyAveragedDynamicSmagorinsky)
//   scale_invariant_operators.jl:10-188  ΣᵢⱼΣᵢⱼᶜᶜᶜ, filter, filtered gradients / strains, ⟨ΣΣᵢⱼ⟩, Σ̄Σ̄ᵢⱼ, Mᵢⱼ, Lᵢⱼ  (ᾱ² = 4, β = 1)
//   smagorinsky.jl:90-104                νₑ = cˢ² Δᶠ² √(2 Σ²)
//
// Five kernels per update_state!: (1) the test-filtered velocities ū, v̄, w̄ once (every filtered gradient reads them, up to one
// cell beyond the interior), (2) Σ = √(ΣᵢⱼΣᵢⱼ) and Σ̄ = √(Σ̄ᵢⱼΣ̄ᵢⱼ) at ccc, halo fill of both, (3) LM = Lᵢⱼ Mᵢⱼ and MM = Mᵢⱼ Mᵢⱼ, (4) their
// averages over the averaging dimensions -- one CTA per output element, fixed summation order, so the result is reproducible --
// (5) νₑ with cˢ² = max(𝒥ᴸᴹ, 𝒥ᴸᴹ_min) / 𝒥ᴹᴹ (𝒥ᴹᴹ > 0).
//
// The pointwise arithmetic is `__host__ __device__` and free of device intrinsics: tests compile this header for the HOST
// (tests/host_dynsmag.cu) and compare it with the numpy restatement of oracle/dynsmag.py without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define OB_DLD(p) __ldg(p)
#else
#define OB_DLD(p) (*(p))
#endif
#define OB_HD __host__ __device__ __forceinline__

namespace ob {

// value at logical (i, j, k) = p[off + i + j*sy + k*sz]  (same convention as Fld<T> of common.cuh)
template <typename T>
struct DField {
    const T *p;
    long off;
    int sy;
    long sz;
    OB_HD T at(int i, int j, int k) const { return OB_DLD(p + (off + i + (long)j * sy + (long)k * sz)); }
};

template <typename T>
struct DynP {
    int N[3], H[3];
    T dx, dy, rdx, rdy, dz, rdz;
    const T *dzc, *rdzc, *rdzf;   // stretched z: per-level Δzᶜ, 1/Δzᶜ, 1/Δzᶠ, pre-offset (logical k); nullptr when regular
    OB_HD T dzC(int k) const { return dzc ? OB_DLD(dzc + k) : dz; }
    OB_HD T rdzC(int k) const { return dzc ? OB_DLD(rdzc + k) : rdz; }
    OB_HD T rdzF(int k) const { return dzc ? OB_DLD(rdzf + k) : rdz; }
    DField<T> u, v, w;            // velocities (halos filled)
    DField<T> ub, vb, wb;         // test-filtered velocities (kernel 1), valid on [2-H, N+H-1]^3
    DField<T> Sg, Sb;             // Σ, Σ̄ (halos filled after kernel 2)
    T *ub_w, *vb_w, *wb_w, *Sg_w, *Sb_w, *LM_w, *MM_w;   // the same arrays / LM, MM for writing (same indexing as the DFields)
    DField<T> LM, MM;
};

// filter (scale_invariant_operators.jl:47-51): (6 f + f(i+1) + f(i-1) + f(j+1) + f(j-1) + f(k+1) + f(k-1)) / 12
template <typename T, class F>
OB_HD T dyn_filter(const F &f, int i, int j, int k) {
    return (6 * f(i, j, k) + f(i + 1, j, k) + f(i - 1, j, k) + f(i, j + 1, k) + f(i, j - 1, k) + f(i, j, k + 1) + f(i, j, k - 1)) / T(12);
}

// The six strain components from the velocity triple (A.u, A.v, A.w): RAW = the velocities, otherwise the filtered ones.
template <typename T, bool RAW>
struct DynStrain {
    const DynP<T> &P;
    OB_HD T U(int i, int j, int k) const { return RAW ? P.u.at(i, j, k) : P.ub.at(i, j, k); }
    OB_HD T V(int i, int j, int k) const { return RAW ? P.v.at(i, j, k) : P.vb.at(i, j, k); }
    OB_HD T W(int i, int j, int k) const { return RAW ? P.w.at(i, j, k) : P.wb.at(i, j, k); }
    OB_HD T s11(int i, int j, int k) const { return (U(i + 1, j, k) - U(i, j, k)) * P.rdx; }                 // ∂xᶜᶜᶜ
    OB_HD T s22(int i, int j, int k) const { return (V(i, j + 1, k) - V(i, j, k)) * P.rdy; }
    OB_HD T s33(int i, int j, int k) const { return (W(i, j, k + 1) - W(i, j, k)) * P.rdzC(k); }
    OB_HD T s12(int i, int j, int k) const {                                                                   // ffc
        return T(0.5) * ((U(i, j, k) - U(i, j - 1, k)) * P.rdy + (V(i, j, k) - V(i - 1, j, k)) * P.rdx);
    }
    OB_HD T s13(int i, int j, int k) const {                                                                   // fcf
        return T(0.5) * ((U(i, j, k) - U(i, j, k - 1)) * P.rdzF(k) + (W(i, j, k) - W(i - 1, j, k)) * P.rdx);
    }
    OB_HD T s23(int i, int j, int k) const {                                                                   // cff
        return T(0.5) * ((V(i, j, k) - V(i, j, k - 1)) * P.rdzF(k) + (W(i, j, k) - W(i, j - 1, k)) * P.rdy);
    }
    // ℑxyᶜᶜᵃ = ℑyᵃᶜᵃ(ℑxᶜᵃᵃ), ℑxzᶜᵃᶜ = ℑzᵃᵃᶜ(ℑxᶜᵃᵃ), ℑyzᵃᶜᶜ = ℑzᵃᵃᶜ(ℑyᵃᶜᵃ) of the ffc / fcf / cff component, optionally squared
    template <bool SQ> OB_HD T q(T x) const { return SQ ? x * x : x; }
    template <bool SQ> OB_HD T Ixy12(int i, int j, int k) const {
        return T(0.5) * (T(0.5) * (q<SQ>(s12(i, j, k)) + q<SQ>(s12(i + 1, j, k))) + T(0.5) * (q<SQ>(s12(i, j + 1, k)) + q<SQ>(s12(i + 1, j + 1, k))));
    }
    template <bool SQ> OB_HD T Ixz13(int i, int j, int k) const {
        return T(0.5) * (T(0.5) * (q<SQ>(s13(i, j, k)) + q<SQ>(s13(i + 1, j, k))) + T(0.5) * (q<SQ>(s13(i, j, k + 1)) + q<SQ>(s13(i + 1, j, k + 1))));
    }
    template <bool SQ> OB_HD T Iyz23(int i, int j, int k) const {
        return T(0.5) * (T(0.5) * (q<SQ>(s23(i, j, k)) + q<SQ>(s23(i, j + 1, k))) + T(0.5) * (q<SQ>(s23(i, j, k + 1)) + q<SQ>(s23(i, j + 1, k + 1))));
    }
    // ΣᵢⱼΣᵢⱼᶜᶜᶜ (scale_invariant_operators.jl:10-13, 112-116)
    OB_HD T double_dot(int i, int j, int k) const {
        const T a = s11(i, j, k), b = s22(i, j, k), c = s33(i, j, k);
        const T tr = a * a + b * b + c * c;
        return tr + 2 * Ixy12<true>(i, j, k) + 2 * Ixz13<true>(i, j, k) + 2 * Iyz23<true>(i, j, k);
    }
    // Σ Σᵢⱼ at ccc, component c = 0..5 (11, 22, 33, 12, 13, 23), with the field S = Σ (RAW) or Σ̄
    OB_HD T SSij(int c, int i, int j, int k) const {
        const T S = RAW ? P.Sg.at(i, j, k) : P.Sb.at(i, j, k);
        const T e = c == 0 ? s11(i, j, k) : c == 1 ? s22(i, j, k) : c == 2 ? s33(i, j, k)
                  : c == 3 ? Ixy12<false>(i, j, k) : c == 4 ? Ixz13<false>(i, j, k) : Iyz23<false>(i, j, k);
        return S * e;
    }
};

// filtered velocities at one point (kernel 1)
template <typename T>
OB_HD void dyn_filter_velocities(const DynP<T> &P, int i, int j, int k, T &ub, T &vb, T &wb) {
    auto U = [&](int a, int b, int c) { return P.u.at(a, b, c); };
    auto V = [&](int a, int b, int c) { return P.v.at(a, b, c); };
    auto W = [&](int a, int b, int c) { return P.w.at(a, b, c); };
    ub = dyn_filter<T>(U, i, j, k);
    vb = dyn_filter<T>(V, i, j, k);
    wb = dyn_filter<T>(W, i, j, k);
}

// Σ and Σ̄ at ccc (kernel 2; _compute_Σ!, _compute_Σ̄!)
template <typename T>
OB_HD void dyn_sigma(const DynP<T> &P, int i, int j, int k, T &Sg, T &Sb) {
    Sg = sqrt(DynStrain<T, true>{P}.double_dot(i, j, k));
    Sb = sqrt(DynStrain<T, false>{P}.double_dot(i, j, k));
}

// LM and MM at ccc (kernel 3; LM_and_MM, dynamic_coefficient.jl:283-304)
template <typename T>
OB_HD void dyn_LM_MM(const DynP<T> &P, int i, int j, int k, T &LM, T &MM) {
    const DynStrain<T, true> R{P};
    const DynStrain<T, false> B{P};
    const T D3 = (P.dx * P.dy) * P.dzC(k);      // volume at ccc
    const T Df = cbrt(D3);
    const T twoD2 = 2 * (Df * Df);
    T M[6], L[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        auto f = [&](int a, int b, int cc) { return R.SSij(c, a, b, cc); };
        const T fSS = dyn_filter<T>(f, i, j, k);             // ⟨ΣΣᵢⱼ⟩
        const T BB = B.SSij(c, i, j, k);                     // Σ̄Σ̄ᵢⱼ
        M[c] = twoD2 * (fSS - T(4) * BB);                    // ᾱ² β = 4
    }
    auto U = [&](int a, int b, int c) { return P.u.at(a, b, c); };
    auto V = [&](int a, int b, int c) { return P.v.at(a, b, c); };
    auto W = [&](int a, int b, int c) { return P.w.at(a, b, c); };
    auto Ub = [&](int a, int b, int c) { return P.ub.at(a, b, c); };
    auto Vb = [&](int a, int b, int c) { return P.vb.at(a, b, c); };
    auto Wb = [&](int a, int b, int c) { return P.wb.at(a, b, c); };
    const T h = T(0.5);
    // uᵢuⱼ at ccc (scale_invariant_operators.jl:155-174)
    auto u1u1 = [&](int a, int b, int c) { const T x = U(a, b, c), y = U(a + 1, b, c); return h * (x * x + y * y); };
    auto u2u2 = [&](int a, int b, int c) { const T x = V(a, b, c), y = V(a, b + 1, c); return h * (x * x + y * y); };
    auto u3u3 = [&](int a, int b, int c) { const T x = W(a, b, c), y = W(a, b, c + 1); return h * (x * x + y * y); };
    auto uc = [&](int a, int b, int c) { return h * (U(a, b, c) + U(a + 1, b, c)); };
    auto vc = [&](int a, int b, int c) { return h * (V(a, b, c) + V(a, b + 1, c)); };
    auto wc = [&](int a, int b, int c) { return h * (W(a, b, c) + W(a, b, c + 1)); };
    auto u1u2 = [&](int a, int b, int c) { return uc(a, b, c) * vc(a, b, c); };
    auto u1u3 = [&](int a, int b, int c) { return uc(a, b, c) * wc(a, b, c); };
    auto u2u3 = [&](int a, int b, int c) { return vc(a, b, c) * wc(a, b, c); };
    const T ub0 = Ub(i, j, k), ub1 = Ub(i + 1, j, k), vb0 = Vb(i, j, k), vb1 = Vb(i, j + 1, k), wb0 = Wb(i, j, k), wb1 = Wb(i, j, k + 1);
    const T ubc = h * (ub0 + ub1), vbc = h * (vb0 + vb1), wbc = h * (wb0 + wb1);
    L[0] = dyn_filter<T>(u1u1, i, j, k) - h * (ub0 * ub0 + ub1 * ub1);
    L[1] = dyn_filter<T>(u2u2, i, j, k) - h * (vb0 * vb0 + vb1 * vb1);
    L[2] = dyn_filter<T>(u3u3, i, j, k) - h * (wb0 * wb0 + wb1 * wb1);
    L[3] = dyn_filter<T>(u1u2, i, j, k) - ubc * vbc;
    L[4] = dyn_filter<T>(u1u3, i, j, k) - ubc * wbc;
    L[5] = dyn_filter<T>(u2u3, i, j, k) - vbc * wbc;
    LM = L[0] * M[0] + L[1] * M[1] + L[2] * M[2] + (2 * L[3]) * M[3] + (2 * L[4]) * M[4] + (2 * L[5]) * M[5];
    MM = M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + (2 * M[3]) * M[3] + (2 * M[4]) * M[4] + (2 * M[5]) * M[5];
}

// νₑ at ccc from the averaged LM, MM (square_smagorinsky_coefficient + _compute_smagorinsky_viscosity!)
template <typename T>
OB_HD T dyn_viscosity(const DynP<T> &P, int i, int j, int k, T JLM, T JMM, T JLM_min) {
    const T num = JLM > JLM_min ? JLM : JLM_min;
    const T cs2 = JMM > 0 ? num / JMM : T(0);
    const T S2 = DynStrain<T, true>{P}.double_dot(i, j, k);
    const T D3 = (P.dx * P.dy) * P.dzC(k);
    const T Df = cbrt(D3);
    return cs2 * (Df * Df) * sqrt(2 * S2);
}

#ifdef __CUDACC__
// ---- kernels -------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) dyn_filter_kernel(const __grid_constant__ DynP<T> P) {
    // cells [2-H, N+H-1] in every direction
    const int ex = P.N[0] + 2 * P.H[0] - 2, ey = P.N[1] + 2 * P.H[1] - 2, ez = P.N[2] + 2 * P.H[2] - 2;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)ex * ey * ez) return;
    const int i = 2 - P.H[0] + (int)(t % ex), j = 2 - P.H[1] + (int)((t / ex) % ey), k = 2 - P.H[2] + (int)(t / ((long)ex * ey));
    T a, b, c;
    dyn_filter_velocities(P, i, j, k, a, b, c);
    P.ub_w[P.ub.off + i + (long)j * P.ub.sy + (long)k * P.ub.sz] = a;
    P.vb_w[P.vb.off + i + (long)j * P.vb.sy + (long)k * P.vb.sz] = b;
    P.wb_w[P.wb.off + i + (long)j * P.wb.sy + (long)k * P.wb.sz] = c;
}
template <typename T>
__device__ __forceinline__ bool dyn_cell(const DynP<T> &P, int &i, int &j, int &k) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)P.N[0] * P.N[1] * P.N[2]) return false;
    i = 1 + (int)(t % P.N[0]); j = 1 + (int)((t / P.N[0]) % P.N[1]); k = 1 + (int)(t / ((long)P.N[0] * P.N[1]));
    return true;
}
template <typename T>
__global__ void __launch_bounds__(128) dyn_sigma_kernel(const __grid_constant__ DynP<T> P) {
    int i, j, k;
    if (!dyn_cell(P, i, j, k)) return;
    T a, b;
    dyn_sigma(P, i, j, k, a, b);
    P.Sg_w[P.Sg.off + i + (long)j * P.Sg.sy + (long)k * P.Sg.sz] = a;
    P.Sb_w[P.Sb.off + i + (long)j * P.Sb.sy + (long)k * P.Sb.sz] = b;
}
template <typename T>
__global__ void __launch_bounds__(128) dyn_lmmm_kernel(const __grid_constant__ DynP<T> P) {
    int i, j, k;
    if (!dyn_cell(P, i, j, k)) return;
    T a, b;
    dyn_LM_MM(P, i, j, k, a, b);
    P.LM_w[P.LM.off + i + (long)j * P.LM.sy + (long)k * P.LM.sz] = a;
    P.MM_w[P.MM.off + i + (long)j * P.MM.sy + (long)k * P.MM.sz] = b;
}
// Average(LM, dims), Average(MM, dims): one CTA per output element (the kept dimensions), threads stride over the averaged
// sub-volume in a fixed order, fixed-shape tree reduction.  avg[d] != 0: dimension d is averaged.  out[0 .. nout) = 𝒥ᴸᴹ,
// out[nout .. 2 nout) = 𝒥ᴹᴹ, output index = (i kept) + nxo * ((j kept) + nyo * (k kept)).
template <typename T>
__global__ void __launch_bounds__(256) dyn_average_kernel(const __grid_constant__ DynP<T> P, int ax, int ay, int az, T *__restrict__ out) {
    const int nxo = ax ? 1 : P.N[0], nyo = ay ? 1 : P.N[1], nzo = az ? 1 : P.N[2];
    const long nout = (long)nxo * nyo * nzo;
    const long o = blockIdx.x;
    const int io = (int)(o % nxo), jo = (int)((o / nxo) % nyo), ko = (int)(o / ((long)nxo * nyo));
    const int mx = ax ? P.N[0] : 1, my = ay ? P.N[1] : 1, mz = az ? P.N[2] : 1;
    const long m = (long)mx * my * mz;
    double s1 = 0, s2 = 0;   // accumulated in Float64 whatever T (the oracle's mean does the same)
    for (long q = threadIdx.x; q < m; q += blockDim.x) {
        const int i = 1 + (ax ? (int)(q % mx) : io), j = 1 + (ay ? (int)((q / mx) % my) : jo), k = 1 + (az ? (int)(q / ((long)mx * my)) : ko);
        s1 += (double)P.LM.at(i, j, k);
        s2 += (double)P.MM.at(i, j, k);
    }
    __shared__ double sh1[256], sh2[256];
    sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { sh1[threadIdx.x] += sh1[threadIdx.x + w]; sh2[threadIdx.x] += sh2[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[o] = (T)(sh1[0] / (double)m); out[nout + o] = (T)(sh2[0] / (double)m); }
}
template <typename T>
__global__ void __launch_bounds__(128) dyn_viscosity_kernel(const __grid_constant__ DynP<T> P, int ax, int ay, int az, const T *__restrict__ J,
                                                            T JLM_min, T *__restrict__ nue, long noff, int nsy, long nsz) {
    int i, j, k;
    if (!dyn_cell(P, i, j, k)) return;
    const int nxo = ax ? 1 : P.N[0], nyo = ay ? 1 : P.N[1], nzo = az ? 1 : P.N[2];
    const long nout = (long)nxo * nyo * nzo;
    const long o = (ax ? 0 : i - 1) + (long)nxo * ((ay ? 0 : j - 1) + (long)nyo * (az ? 0 : k - 1));
    nue[noff + i + (long)j * nsy + (long)k * nsz] = dyn_viscosity(P, i, j, k, J[o], J[nout + o], JLM_min);
}
#endif  // __CUDACC__

}  // namespace ob
