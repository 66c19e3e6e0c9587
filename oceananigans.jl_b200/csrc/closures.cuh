// closures.cuh -- eddy-viscosity / eddy-diffusivity closure fields (reference kernels K9 of SURVEY.md §2a):
//   _compute_smagorinsky_viscosity!   Smagorinskys/smagorinsky.jl:90-104, lilly_coefficient.jl:129-142,
//                                     scale_invariant_operators.jl:10-13
//   _compute_AMD_viscosity! / _compute_AMD_diffusivity!
//                                     anisotropic_minimum_dissipation.jl:161-204 and the "30 terms" :251-358,
//                                     normalised gradients velocity_tracer_gradients.jl:126-260
// One thread per cell over the interior (:xyz); the AMD kernel computes νₑ (blockIdx.y == 0) and every κₑ
// (blockIdx.y == 1 + tracer) in ONE launch (the reference launches 1 + n_tracers kernels).
#pragma once
#include "tendency.cuh"

namespace ob {

// ℑ to centre along D of a pointwise functor: 0.5 * (f(i) + f(i+1)); Flat => identity (interpolation_operators.jl:8-28)
template <typename T, int D, class F>
__device__ __forceinline__ T Ic1(const GridD<T> &g, const F &f, int i, int j, int k) {
    if (g.topo[D] == FLAT) return f(i, j, k);
    int a = i, b = j, c = k;
    shift<D>(a, b, c, 1);
    return T(0.5) * (f(i, j, k) + f(a, b, c));
}
// ℑ_{D2}ᶜ(ℑ_{D1}ᶜ f): ℑxyᶜᶜᵃ = <1,0>, ℑxzᶜᵃᶜ = <2,0>, ℑyzᵃᶜᶜ = <2,1>   (interpolation_operators.jl:45-53)
template <typename T, int D2, int D1, class F>
__device__ __forceinline__ T Ic2(const GridD<T> &g, const F &f, int i, int j, int k) {
    auto inner = [&](int a, int b, int c) { return Ic1<T, D1>(g, f, a, b, c); };
    return Ic1<T, D2>(g, inner, i, j, k);
}

template <typename T> struct ClosP {
    TendP<T> P;
};

template <typename T>
__global__ void __launch_bounds__(128) smagorinsky_kernel(const __grid_constant__ TendP<T> P, int m) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    const GridD<T> &g = P.g;
    Grad<T> G{P};
    const T s11 = G.S11(i, j, k), s22 = G.S22(i, j, k), s33 = G.S33(i, j, k);
    const T tr = s11 * s11 + s22 * s22 + s33 * s33;
    auto s12sq = [&](int a, int b, int c) { T s = G.S12(a, b, c); return s * s; };
    auto s13sq = [&](int a, int b, int c) { T s = G.S13(a, b, c); return s * s; };
    auto s23sq = [&](int a, int b, int c) { T s = G.S23(a, b, c); return s * s; };
    const T SS = tr + 2 * Ic2<T, 1, 0>(g, s12sq, i, j, k) + 2 * Ic2<T, 2, 0>(g, s13sq, i, j, k) + 2 * Ic2<T, 2, 1>(g, s23sq, i, j, k);
    const T D3 = g.dx * g.dy * g.dzC(k);
    const T Df = cbrt(D3);
    const ClosureD<T> &c = P.cl[m];
    T cs2;
    if (c.lilly) {
        auto dzb = [&](int a, int b, int cc) {
            return (g.topo[2] == FLAT ? T(0) : bpert(P, a, b, cc) - bpert(P, a, b, cc - 1)) * g.rdzF(cc);
        };
        const T N2 = Ic1<T, 2>(g, dzb, i, j, k);
        const T N2p = fmax(T(0), N2);
        const T s2 = T(1) - fmin(T(1), c.cb * N2p / SS);
        const T st = (SS == 0) ? T(0) : sqrt(s2);
        cs2 = st * (c.cs * c.cs);
    } else {
        cs2 = c.cs * c.cs;
    }
    P.nue[m](i, j, k) = cs2 * (Df * Df) * sqrt(2 * SS);
}

template <typename T>
__global__ void __launch_bounds__(128) amd_kernel(const __grid_constant__ TendP<T> P, int m) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    const GridD<T> &g = P.g;
    Grad<T> G{P};
    const ClosureD<T> &cl = P.cl[m];
    // filter widths: 2Δ at ccc, evaluated at the stencil point's own index whatever its location (:231-243)
    auto dfx = [&](int) { return 2 * g.dx; };
    auto dfy = [&](int) { return 2 * g.dy; };
    auto dfz = [&](int c) { return 2 * g.dzC(c); };
    auto n_dx_v = [&](int a, int b, int c) { return dfx(a) / dfy(b) * G.dx_v(a, b, c); };
    auto n_dy_u = [&](int a, int b, int c) { return dfy(b) / dfx(a) * G.dy_u(a, b, c); };
    auto n_dx_w = [&](int a, int b, int c) { return dfx(a) / dfz(c) * G.dx_w(a, b, c); };
    auto n_dz_u = [&](int a, int b, int c) { return dfz(c) / dfx(a) * G.dz_u(a, b, c); };
    auto n_dy_w = [&](int a, int b, int c) { return dfy(b) / dfz(c) * G.dy_w(a, b, c); };
    auto n_dz_v = [&](int a, int b, int c) { return dfz(c) / dfy(b) * G.dz_v(a, b, c); };
    const T fx = dfx(i), fy = dfy(j), fz = dfz(k);
    const T delta2 = 3 / (1 / (fx * fx) + 1 / (fy * fy) + 1 / (fz * fz));
    const T ux = G.dx_u(i, j, k), vy = G.dy_v(i, j, k), wz = G.dz_w(i, j, k);
#define IXY(f) Ic2<T, 1, 0>(g, f, i, j, k)
#define IXZ(f) Ic2<T, 2, 0>(g, f, i, j, k)
#define IYZ(f) Ic2<T, 2, 1>(g, f, i, j, k)
    if (blockIdx.y == 0) {
        auto n_S12 = [&](int a, int b, int c) { return T(0.5) * (n_dy_u(a, b, c) + n_dx_v(a, b, c)); };
        auto n_S13 = [&](int a, int b, int c) { return T(0.5) * (n_dz_u(a, b, c) + n_dx_w(a, b, c)); };
        auto n_S23 = [&](int a, int b, int c) { return T(0.5) * (n_dz_v(a, b, c) + n_dy_w(a, b, c)); };
        auto sq = [](T x) { return x * x; };
        auto n_dx_v2 = [&](int a, int b, int c) { return sq(n_dx_v(a, b, c)); };
        auto n_dy_u2 = [&](int a, int b, int c) { return sq(n_dy_u(a, b, c)); };
        auto n_dx_w2 = [&](int a, int b, int c) { return sq(n_dx_w(a, b, c)); };
        auto n_dz_u2 = [&](int a, int b, int c) { return sq(n_dz_u(a, b, c)); };
        auto n_dy_w2 = [&](int a, int b, int c) { return sq(n_dy_w(a, b, c)); };
        auto n_dz_v2 = [&](int a, int b, int c) { return sq(n_dz_v(a, b, c)); };
        const T q = ux * ux + vy * vy + wz * wz + IXY(n_dx_v2) + IXY(n_dy_u2) + IXZ(n_dx_w2) + IXZ(n_dz_u2) + IYZ(n_dy_w2) + IYZ(n_dz_v2);
        T nu = 0;
        if (q != 0) {
            auto n_dx_v_S12 = [&](int a, int b, int c) { return n_dx_v(a, b, c) * n_S12(a, b, c); };
            auto n_dy_u_S12 = [&](int a, int b, int c) { return n_dy_u(a, b, c) * n_S12(a, b, c); };
            auto n_dx_w_S13 = [&](int a, int b, int c) { return n_dx_w(a, b, c) * n_S13(a, b, c); };
            auto n_dz_u_S13 = [&](int a, int b, int c) { return n_dz_u(a, b, c) * n_S13(a, b, c); };
            auto n_dz_v_S23 = [&](int a, int b, int c) { return n_dz_v(a, b, c) * n_S23(a, b, c); };
            auto n_dy_w_S23 = [&](int a, int b, int c) { return n_dy_w(a, b, c) * n_S23(a, b, c); };
            const T r1 = ux * (ux * ux) + vy * IXY(n_dx_v2) + wz * IXZ(n_dx_w2) + 2 * ux * IXY(n_dx_v_S12) + 2 * ux * IXZ(n_dx_w_S13) +
                         2 * IXY(n_dx_v) * IXZ(n_dx_w) * IYZ(n_S23);
            const T r2 = ux * IXY(n_dy_u2) + vy * (vy * vy) + wz * IYZ(n_dy_w2) + 2 * vy * IXY(n_dy_u_S12) +
                         2 * IXY(n_dy_u) * IYZ(n_dy_w) * IXZ(n_S13) + 2 * vy * IYZ(n_dy_w_S23);
            const T r3 = ux * IXZ(n_dz_u2) + vy * IYZ(n_dz_v2) + wz * (wz * wz) + 2 * IXZ(n_dz_u) * IYZ(n_dz_v) * IXY(n_S12) +
                         2 * wz * IXZ(n_dz_u_S13) + 2 * wz * IYZ(n_dz_v_S23);
            const T r = r1 + r2 + r3;
            T cbz = 0;
            if (cl.amd_has_cb) {
                auto dxb = [&](int a, int b, int c) { return (g.topo[0] == FLAT ? T(0) : bpert(P, a, b, c) - bpert(P, a - 1, b, c)) * g.rdx; };
                auto dyb = [&](int a, int b, int c) { return (g.topo[1] == FLAT ? T(0) : bpert(P, a, b, c) - bpert(P, a, b - 1, c)) * g.rdy; };
                auto dzb = [&](int a, int b, int c) { return (g.topo[2] == FLAT ? T(0) : bpert(P, a, b, c) - bpert(P, a, b, c - 1)) * g.rdzF(c); };
                const T wxbx = IXZ(n_dx_w) * fx * Ic1<T, 0>(g, dxb, i, j, k);
                const T wyby = IYZ(n_dy_w) * fy * Ic1<T, 1>(g, dyb, i, j, k);
                const T wzbz = wz * fz * Ic1<T, 2>(g, dzb, i, j, k);
                cbz = cl.cb * (wxbx + wyby + wzbz);
            }
            cbz = cbz / fz;
            nu = -cl.Cnu * delta2 * (r - cbz) / q;
        }
        P.nue[m](i, j, k) = fmax(T(0), nu);
    } else {
        const int t = blockIdx.y - 1;
        const Fld<T> &c = P.c[t];
        auto n_dx_c = [&](int a, int b, int cc) { return dfx(a) * ((g.topo[0] == FLAT ? T(0) : c.ld(a, b, cc) - c.ld(a - 1, b, cc)) * g.rdx); };
        auto n_dy_c = [&](int a, int b, int cc) { return dfy(b) * ((g.topo[1] == FLAT ? T(0) : c.ld(a, b, cc) - c.ld(a, b - 1, cc)) * g.rdy); };
        auto n_dz_c = [&](int a, int b, int cc) { return dfz(cc) * ((g.topo[2] == FLAT ? T(0) : c.ld(a, b, cc) - c.ld(a, b, cc - 1)) * g.rdzF(cc)); };
        auto n_dx_c2 = [&](int a, int b, int cc) { T x = n_dx_c(a, b, cc); return x * x; };
        auto n_dy_c2 = [&](int a, int b, int cc) { T x = n_dy_c(a, b, cc); return x * x; };
        auto n_dz_c2 = [&](int a, int b, int cc) { T x = n_dz_c(a, b, cc); return x * x; };
        const T icx2 = Ic1<T, 0>(g, n_dx_c2, i, j, k), icy2 = Ic1<T, 1>(g, n_dy_c2, i, j, k), icz2 = Ic1<T, 2>(g, n_dz_c2, i, j, k);
        const T sigma = icx2 + icy2 + icz2;
        T kap_ = 0;
        if (sigma != 0) {
            const T icx = Ic1<T, 0>(g, n_dx_c, i, j, k), icy = Ic1<T, 1>(g, n_dy_c, i, j, k), icz = Ic1<T, 2>(g, n_dz_c, i, j, k);
            // cy_uy uses ℑxzᶜᵃᶜ of norm_∂y_w exactly as the reference does (:345)
            const T cx = ux * icx2 + IXY(n_dx_v) * icx * icy + IXZ(n_dx_w) * icx * icz;
            const T cy = IXY(n_dy_u) * icy * icx + vy * icy2 + IXZ(n_dy_w) * icy * icz;
            const T cz = IXZ(n_dz_u) * icz * icx + IYZ(n_dz_v) * icz * icy + wz * icz2;
            const T theta = cx + cy + cz;
            kap_ = -cl.Ckappa[t] * delta2 * theta / sigma;
        }
        P.kappae[m][t](i, j, k) = fmax(T(0), kap_);
    }
#undef IXY
#undef IXZ
#undef IYZ
}

}  // namespace ob
