// stencils.cuh -- device stencil algebra: Centered / Z-WENO reconstruction with the reference's Bounded-direction
// fallback chain, advective fluxes, strain rates, closure fluxes.  Everything is compile-time unrolled on the
// scheme (kind, buffer); topology is a launch-uniform runtime branch.
//
// Reference semantics restated here (file:line under /root/reference/src):
//   Advection/weno_interpolants.jl:71-562, reconstruction_coefficients.jl:62-165, centered_reconstruction.jl:54-63,
//   upwind_biased_reconstruction.jl:89-94, topologically_conditional_interpolation.jl:43-128,
//   upwind_biased_advective_fluxes.jl:11-121, centered_advective_fluxes.jl:19-37, flat_advective_fluxes.jl:9-49,
//   TurbulenceClosures/closure_kernel_operators.jl:20-46, abstract_scalar_diffusivity_closure.jl:209-262,330-351,
//   velocity_tracer_gradients.jl:6-42, Operators/*.jl
#pragma once
#include "common.cuh"
#include "coef_literals.h"

namespace ob {

// ---- coefficient tables (filled by the host exactly as the reference computes them) -------------------------
struct CoefTables64 {
    double weno_coeff[OB_MAXBUF + 1][OB_MAXBUF][OB_MAXBUF];
    double weno_beta[OB_MAXBUF + 1][OB_MAXBUF][21];
    double weno_cstar[OB_MAXBUF + 1][OB_MAXBUF];
    double cen_coeff[OB_MAXBUF + 1][2 * OB_MAXBUF];
    double weno_eps;
};
struct CoefTables32 {
    float weno_coeff[OB_MAXBUF + 1][OB_MAXBUF][OB_MAXBUF];
    float weno_beta[OB_MAXBUF + 1][OB_MAXBUF][21];
    float weno_cstar[OB_MAXBUF + 1][OB_MAXBUF];
    float cen_coeff[OB_MAXBUF + 1][2 * OB_MAXBUF];
    float weno_eps;
};
static __constant__ CoefTables64 c_tab64;
static __constant__ CoefTables32 c_tab32;
template <typename T> struct Tab;
template <> struct Tab<double> { static __device__ __forceinline__ const CoefTables64 &get() { return c_tab64; } };
template <> struct Tab<float> { static __device__ __forceinline__ const CoefTables32 &get() { return c_tab32; } };

template <typename T> __device__ __forceinline__ T fma_(T a, T b, T c);
template <> __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }

// Fast reciprocal: rcp.approx + one cubic Newton step, the construction the reference uses on CUDA devices
// (ext/OceananigansCUDAExt.jl:147-163).
__device__ __forceinline__ double fast_rcp(double b) {
    double inv;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(b));
    double e = fma(inv, -b, 1.0);
    e = fma(e, e, e);
    return fma(e, inv, inv);
}
__device__ __forceinline__ float fast_rcp(float b) {
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(b));
    float e = fmaf(inv, -b, 1.0f);
    e = fmaf(e, e, e);
    return fmaf(e, inv, inv);
}
template <bool FAST, typename T> __device__ __forceinline__ T div_(T a, T b) {
    if constexpr (FAST) return a * fast_rcp(b);
    else return a / b;
}

// ---- schemes -------------------------------------------------------------------------------------------------
template <int KIND, int NB> struct Scheme { static constexpr int kind = KIND; static constexpr int n = NB; };
template <class S> struct BufferOf { using type = Scheme<ADV_CENTERED, (S::n > 1 ? S::n - 1 : 1)>; };
// WENO(order) -> WENO(order-2) ... WENO(3) -> Centered(2)   (weno_reconstruction.jl:126-133)
template <int NB> struct BufferOf<Scheme<ADV_WENO, NB>> {
    using type = typename std::conditional<(NB > 2), Scheme<ADV_WENO, NB - 1>, Scheme<ADV_CENTERED, 1>>::type;
};
template <class S> struct AdvectingOf { using type = S; };
template <int NB> struct AdvectingOf<Scheme<ADV_WENO, NB>> { using type = Scheme<ADV_CENTERED, NB - 1>; };

template <int DIR> __device__ __forceinline__ void shift(int &i, int &j, int &k, int s) {
    if constexpr (DIR == 0) i += s; else if constexpr (DIR == 1) j += s; else k += s;
}

// symmetric_interpolate_*ᶠ for Centered{N}: @muladd Σ C_m ψ[idx + m - N - 1]
template <typename T, int N, int DIR, class G>
__device__ __forceinline__ T centered_face(const G &get, int i, int j, int k) {
    const auto &tab = Tab<T>::get();
    T acc;
#pragma unroll
    for (int m = 1; m <= 2 * N; m++) {
        int ii = i, jj = j, kk = k;
        shift<DIR>(ii, jj, kk, m - N - 1);
        T v = get(ii, jj, kk);
        acc = (m == 1) ? tab.cen_coeff[N][0] * v : fma_(tab.cen_coeff[N][m - 1], v, acc);
    }
    return acc;
}

// Z-WENO weights and reconstruction from the 2N-1 upwind-ordered values v[0..2N-2]
// (v[m] = ψ[face - N + m] for LeftBias, ψ[face + N - 1 - m] for RightBias; the right-biased sub-stencils of
// weno_interpolants.jl:448-471 are the left-biased ones on the mirrored stencil).
// LIT: every coefficient is a compile-time literal (coef_literals.h, the same values as the constant-bank tables) -- the
// operations and their order are unchanged, only the operand kind differs (immediate / literal pool instead of LDC).
template <typename T, int N, bool FAST, bool LIT = false>
__device__ __forceinline__ T weno_from_values(const T (&v)[2 * N - 1]) {
    const auto &tab = Tab<T>::get();
    // smoothness coefficients (weno_interpolants.jl:169-192) come from the constant bank
    auto wb = [&](int r, int c) -> T {
        if constexpr (LIT) return CoefLit<T, N>::beta(r, c);
        else
#ifdef OB_BETA_LITERALS  // OFF: literal operands let the compiler re-associate the ill-conditioned quadratic form (measured: 1e-8 on a tracer with a large mean)
        if constexpr (N == 3) {
            constexpr T tbl[3][6] = {{10, -31, 11, 25, -19, 4}, {4, -13, 5, 13, -13, 4}, {4, -19, 11, 25, -31, 10}};
            return tbl[r][c];
        } else if constexpr (N == 2) {
            constexpr T tbl[3] = {1, -2, 1};
            return tbl[c];
        } else
#endif
        {
            return tab.weno_beta[N][r][c];
        }
    };
    T beta[N];
#pragma unroll
    for (int r = 0; r < N; r++) {
        T ps[N];
#pragma unroll
        for (int q = 0; q < N; q++) {
            if constexpr (sizeof(T) == 4) ps[q] = v[N - 1 - r + q] - v[N - 1 - r + N / 2];  // weno_interpolants.jl:272-280
            else ps[q] = v[N - 1 - r + q];
        }
        int c = 0;
        T b = 0;
#pragma unroll
        for (int s = 0; s < N - 1; s++) {
            T inner = wb(r, c) * ps[s];
#pragma unroll
            for (int q = s + 1; q < N; q++) inner = fma_(wb(r, c + q - s), ps[q], inner);
            b = (s == 0) ? ps[s] * inner : fma_(ps[s], inner, b);
            c += N - s;
        }
        b = fma_(ps[N - 1] * ps[N - 1], wb(r, c), b);
        beta[r] = b;
    }
    T tau;
    if constexpr (N == 2) tau = fabs(beta[0] - beta[1]);
    else if constexpr (N == 3) tau = fabs(beta[0] - beta[2]);
    else if constexpr (N == 4) tau = fabs(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]);
    else if constexpr (N == 5) tau = fabs(beta[0] + 2 * beta[1] - 6 * beta[2] + 2 * beta[3] + beta[4]);
    else tau = fabs(beta[0] + 36 * beta[1] + 135 * beta[2] - 135 * beta[3] - 36 * beta[4] - beta[5]);
    T alpha[N];
    T sum = 0;
#pragma unroll
    for (int r = 0; r < N; r++) {
        T q, cs;
        if constexpr (LIT) { q = div_<FAST>(tau, beta[r] + CoefLit<T, N>::eps()); cs = CoefLit<T, N>::cstar(r); }
        else { q = div_<FAST>(tau, beta[r] + tab.weno_eps); cs = tab.weno_cstar[N][r]; }
        alpha[r] = cs * (1 + q * q);
        sum = (r == 0) ? alpha[r] : sum + alpha[r];
    }
    T inv = FAST ? fast_rcp(sum) : 1 / sum;
    T res = 0;
#pragma unroll
    for (int r = 0; r < N; r++) {
        auto wc = [&](int q) -> T { if constexpr (LIT) return CoefLit<T, N>::coeff(r, q); else return tab.weno_coeff[N][r][q]; };
        T p = wc(0) * v[N - 1 - r];
#pragma unroll
        for (int q = 1; q < N; q++) p = p + wc(q) * v[N - 1 - r + q];
        res = (r == 0) ? (alpha[r] * inv) * p : fma_(alpha[r] * inv, p, res);
    }
    return res;
}

template <typename T, int N, int DIR, bool FAST, class G>
__device__ __forceinline__ T weno_face(const G &get, bool left, int i, int j, int k) {
    T v[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) {
        int ii = i, jj = j, kk = k;
        shift<DIR>(ii, jj, kk, left ? (m - N) : (N - 1 - m));
        v[m] = get(ii, jj, kk);
    }
    return weno_from_values<T, N, FAST>(v);
}

// scheme-level plain interpolations (no topology logic)
template <typename T, class S, int DIR, bool CENTER, class G>
__device__ __forceinline__ T sym_plain(const G &get, int i, int j, int k) {
    using C = typename AdvectingOf<S>::type;
    if constexpr (CENTER) shift<DIR>(i, j, k, 1);
    return centered_face<T, C::n, DIR>(get, i, j, k);
}
template <typename T, class S, int DIR, bool CENTER, bool FAST, class G>
__device__ __forceinline__ T biased_plain(const G &get, bool left, int i, int j, int k) {
    if constexpr (CENTER) shift<DIR>(i, j, k, 1);
    if constexpr (S::kind == ADV_WENO) return weno_face<T, S::n, DIR, FAST>(get, left, i, j, k);
    else return centered_face<T, S::n, DIR>(get, i, j, k);
}

// outside_*_halo predicates (topologically_conditional_interpolation.jl:52-58)
template <bool CENTER> __device__ __forceinline__ bool outside_sym(int i, int N, int H) {
    return CENTER ? ((i >= H) & (i <= N + 1 - H)) : ((i >= H + 1) & (i <= N + 1 - H));
}
template <bool CENTER> __device__ __forceinline__ bool outside_biased(int i, int N, int H) {
    return CENTER ? ((i >= H) & (i <= N + 1 - (H - 1)) & (i >= H - 1) & (i <= N + 1 - H))
                  : ((i >= H + 1) & (i <= N + 1 - (H - 1)) & (i >= H) & (i <= N + 1 - H));
}

template <typename T, class S, int DIR, bool CENTER, class G>
__device__ __forceinline__ T sym_bounded(const G &get, int idx, int N, int i, int j, int k) {
    if constexpr (S::kind == ADV_CENTERED && S::n == 1) return sym_plain<T, S, DIR, CENTER>(get, i, j, k);
    else {
        if (outside_sym<CENTER>(idx, N, S::n)) return sym_plain<T, S, DIR, CENTER>(get, i, j, k);
        return sym_bounded<T, typename BufferOf<S>::type, DIR, CENTER>(get, idx, N, i, j, k);
    }
}
template <typename T, class S, int DIR, bool CENTER, bool FAST, class G>
__device__ __forceinline__ T biased_bounded(const G &get, bool left, int idx, int N, int i, int j, int k) {
    if constexpr (S::kind == ADV_CENTERED && S::n == 1) return biased_plain<T, S, DIR, CENTER, FAST>(get, left, i, j, k);
    else {
        if (outside_biased<CENTER>(idx, N, S::n)) return biased_plain<T, S, DIR, CENTER, FAST>(get, left, i, j, k);
        return biased_bounded<T, typename BufferOf<S>::type, DIR, CENTER, FAST>(get, left, idx, N, i, j, k);
    }
}

// _symmetric_interpolate_* / _biased_interpolate_*
template <typename T, class S, int DIR, bool CENTER, class G>
__device__ __forceinline__ T sym_interp(const GridD<T> &g, const G &get, int i, int j, int k) {
    const int topo = g.topo[DIR];
    if (topo == FLAT) return get(i, j, k);
    if (topo == BOUNDED) return sym_bounded<T, S, DIR, CENTER>(get, DIR == 0 ? i : DIR == 1 ? j : k, g.N[DIR], i, j, k);
    return sym_plain<T, S, DIR, CENTER>(get, i, j, k);
}
template <typename T, class S, int DIR, bool CENTER, bool FAST, class G>
__device__ __forceinline__ T biased_interp(const GridD<T> &g, const G &get, bool left, int i, int j, int k) {
    const int topo = g.topo[DIR];
    if (topo == FLAT) return get(i, j, k);
    if (topo == BOUNDED)
        return biased_bounded<T, S, DIR, CENTER, FAST>(get, left, DIR == 0 ? i : DIR == 1 ? j : k, g.N[DIR], i, j, k);
    return biased_plain<T, S, DIR, CENTER, FAST>(get, left, i, j, k);
}

// ---- getters -----------------------------------------------------------------------------------------------------
template <typename T> struct GetF {
    const Fld<T> &f;
    __device__ __forceinline__ T operator()(int i, int j, int k) const { return f.ld(i, j, k); }
};
// Ax_qᶠᶜᶜ(U), Ay_qᶜᶠᶜ(V), Az_qᶜᶜᶠ(W)  (products_between_fields_and_grid_metrics.jl:5-14)
template <typename T, int DIR> struct GetAq {
    const Fld<T> &f;
    const GridD<T> &g;
    __device__ __forceinline__ T operator()(int i, int j, int k) const {
        T A;
        if constexpr (DIR == 0) A = g.dy * g.dzC(k);
        else if constexpr (DIR == 1) A = g.dx * g.dzC(k);
        else A = g.dx * g.dy;
        return A * f.ld(i, j, k);
    }
};

}  // namespace ob
