// tendency.cuh -- fused momentum + tracer tendency kernels (reference kernels K1+K2 of SURVEY.md §2a:
// compute_Gu!/compute_Gv!/compute_Gw!/compute_Gc!, compute_nonhydrostatic_tendencies.jl:107-150, which the
// reference launches as 3+n separate kernels).  ONE launch here writes all 3+n tendencies.
//
// `tendency_generic_kernel`: one thread per cell, any topology / scheme / closure combination.
// `tendency_tiled_kernel` (tendency_tiled.cuh): shared-memory tiled fast path for the headline configs.
#pragma once
#include "stencils.cuh"

namespace ob {

// x extent of the cells one CTA of the marching kernel produces (32 lanes, the last one is the flux-overlap lane); a
// launch can be restricted to the x tiles [tx_lo, tx_hi) -- the distributed model computes the tiles that do not read
// x halos while the halo exchange is in flight (interleave_communication_and_computation.jl:36-74)
#define OB_TILE_X 31

template <typename T>
struct TendP {
    GridD<T> g;
    Fld<T> u, v, w, c[OB_MAXTR];
    Fld<T> Gu, Gv, Gw, Gc[OB_MAXTR];
    Fld<T> pHY;
    Fld<T> nue[OB_MAXCL];
    Fld<T> kappae[OB_MAXCL][OB_MAXTR];
    ClosureD<T> cl[OB_MAXCL];
    int ntr, ncl, has_pHY;
    int buoy, ib, iT, iS;
    T grav, alpha, beta;
    int has_cor;
    T f;
};

#define DXF (P.g.dx)
#define DXC (P.g.dx)
#define DYF (P.g.dy)
#define DYC (P.g.dy)
#define DZF(k) (P.g.dzF(k))
#define DZC(k) (P.g.dzC(k))

// ---- advective fluxes ---------------------------------------------------------------------------------------
// ADV: direction of the advecting velocity (0 U, 1 V, 2 W); COMP: advected component (0 u, 1 v, 2 w)
template <typename T, class S, bool FAST, int ADV, int COMP>
__device__ __forceinline__ T mom_flux(const TendP<T> &P, int i, int j, int k) {
    if constexpr (S::kind == ADV_NONE) return 0;
    else {
        if (P.g.topo[ADV] == FLAT) return 0;
        const Fld<T> &Uf = ADV == 0 ? P.u : ADV == 1 ? P.v : P.w;
        const Fld<T> &q = COMP == 0 ? P.u : COMP == 1 ? P.v : P.w;
        constexpr bool CEN = (ADV == COMP);
        if constexpr (S::kind == ADV_WENO) {
            T ut = sym_interp<T, S, COMP, CEN>(P.g, GetAq<T, ADV>{Uf, P.g}, i, j, k);
            T qr = biased_interp<T, S, ADV, CEN, FAST>(P.g, GetF<T>{q}, ut > 0, i, j, k);
            return ut * qr;
        } else {
            T A;
            if constexpr (ADV == 0) A = (COMP == 1 ? DYF : DYC) * (COMP == 2 ? DZF(k) : DZC(k));
            else if constexpr (ADV == 1) A = (COMP == 0 ? DXF : DXC) * (COMP == 2 ? DZF(k) : DZC(k));
            else A = (COMP == 0 ? DXF : DXC) * (COMP == 1 ? DYF : DYC);
            T ut = sym_interp<T, S, COMP, CEN>(P.g, GetF<T>{Uf}, i, j, k);
            T qr = sym_interp<T, S, ADV, CEN>(P.g, GetF<T>{q}, i, j, k);
            return A * ut * qr;
        }
    }
}

template <typename T, class S, bool FAST, int DIR>
__device__ __forceinline__ T tracer_flux(const TendP<T> &P, const Fld<T> &c, int i, int j, int k) {
    if constexpr (S::kind == ADV_NONE) return 0;
    else {
        if (P.g.topo[DIR] == FLAT) return 0;
        const Fld<T> &Uf = DIR == 0 ? P.u : DIR == 1 ? P.v : P.w;
        T A = DIR == 0 ? DYC * DZC(k) : DIR == 1 ? DXC * DZC(k) : DXC * DYC;
        T ut = Uf.ld(i, j, k);
        if constexpr (S::kind == ADV_WENO) {
            T cr = biased_interp<T, S, DIR, false, FAST>(P.g, GetF<T>{c}, ut > 0, i, j, k);
            return A * ut * cr;
        } else {
            return (A * ut) * sym_interp<T, S, DIR, false>(P.g, GetF<T>{c}, i, j, k);
        }
    }
}

// ---- gradients / strain (velocity_tracer_gradients.jl:6-42); δ in a Flat direction is 0 --------------------------
template <typename T> struct Grad {
    const TendP<T> &P;
    __device__ __forceinline__ bool fx() const { return P.g.topo[0] == FLAT; }
    __device__ __forceinline__ bool fy() const { return P.g.topo[1] == FLAT; }
    __device__ __forceinline__ bool fz() const { return P.g.topo[2] == FLAT; }
    __device__ __forceinline__ T dx_u(int i, int j, int k) const { return (fx() ? T(0) : P.u.ld(i + 1, j, k) - P.u.ld(i, j, k)) * P.g.rdx; }
    __device__ __forceinline__ T dy_v(int i, int j, int k) const { return (fy() ? T(0) : P.v.ld(i, j + 1, k) - P.v.ld(i, j, k)) * P.g.rdy; }
    __device__ __forceinline__ T dz_w(int i, int j, int k) const { return (fz() ? T(0) : P.w.ld(i, j, k + 1) - P.w.ld(i, j, k)) * P.g.rdzC(k); }
    __device__ __forceinline__ T dx_v(int i, int j, int k) const { return (fx() ? T(0) : P.v.ld(i, j, k) - P.v.ld(i - 1, j, k)) * P.g.rdx; }
    __device__ __forceinline__ T dy_u(int i, int j, int k) const { return (fy() ? T(0) : P.u.ld(i, j, k) - P.u.ld(i, j - 1, k)) * P.g.rdy; }
    __device__ __forceinline__ T dx_w(int i, int j, int k) const { return (fx() ? T(0) : P.w.ld(i, j, k) - P.w.ld(i - 1, j, k)) * P.g.rdx; }
    __device__ __forceinline__ T dz_u(int i, int j, int k) const { return (fz() ? T(0) : P.u.ld(i, j, k) - P.u.ld(i, j, k - 1)) * P.g.rdzF(k); }
    __device__ __forceinline__ T dy_w(int i, int j, int k) const { return (fy() ? T(0) : P.w.ld(i, j, k) - P.w.ld(i, j - 1, k)) * P.g.rdy; }
    __device__ __forceinline__ T dz_v(int i, int j, int k) const { return (fz() ? T(0) : P.v.ld(i, j, k) - P.v.ld(i, j, k - 1)) * P.g.rdzF(k); }
    __device__ __forceinline__ T S11(int i, int j, int k) const { return dx_u(i, j, k); }
    __device__ __forceinline__ T S22(int i, int j, int k) const { return dy_v(i, j, k); }
    __device__ __forceinline__ T S33(int i, int j, int k) const { return dz_w(i, j, k); }
    __device__ __forceinline__ T S12(int i, int j, int k) const { return T(0.5) * add_rn(dy_u(i, j, k), dx_v(i, j, k)); }
    __device__ __forceinline__ T S13(int i, int j, int k) const { return T(0.5) * add_rn(dz_u(i, j, k), dx_w(i, j, k)); }
    __device__ __forceinline__ T S23(int i, int j, int k) const { return T(0.5) * add_rn(dz_v(i, j, k), dy_w(i, j, k)); }
};

// two-point interpolation of a ccc array to faces (interpolation_operators.jl:8-71), Flat => identity
template <typename T, int D> __device__ __forceinline__ T If1(const GridD<T> &g, const Fld<T> &f, int i, int j, int k) {
    if (g.topo[D] == FLAT) return f.ld(i, j, k);
    int a = i, b = j, c = k;
    shift<D>(a, b, c, -1);
    return T(0.5) * (f.ld(a, b, c) + f.ld(i, j, k));
}
// ℑ_{D2}ᶠ(ℑ_{D1}ᶠ f): e.g. ℑxyᶠᶠᵃ = ℑyᵃᶠᵃ(ℑxᶠᵃᵃ f)
template <typename T, int D2, int D1> __device__ __forceinline__ T If2(const GridD<T> &g, const Fld<T> &f, int i, int j, int k) {
    if (g.topo[D2] == FLAT) return If1<T, D1>(g, f, i, j, k);
    int a = i, b = j, c = k;
    shift<D2>(a, b, c, -1);
    return T(0.5) * (If1<T, D1>(g, f, a, b, c) + If1<T, D1>(g, f, i, j, k));
}

template <typename T> struct Visc {
    const TendP<T> &P;
    int m;
    __device__ __forceinline__ bool is_const() const { return P.cl[m].kind == CL_SCALAR; }
    __device__ __forceinline__ T ccc(int i, int j, int k) const { return is_const() ? P.cl[m].nu : P.nue[m].ld(i, j, k); }
    __device__ __forceinline__ T ffc(int i, int j, int k) const { return is_const() ? P.cl[m].nu : If2<T, 1, 0>(P.g, P.nue[m], i, j, k); }
    __device__ __forceinline__ T fcf(int i, int j, int k) const { return is_const() ? P.cl[m].nu : If2<T, 2, 0>(P.g, P.nue[m], i, j, k); }
    __device__ __forceinline__ T cff(int i, int j, int k) const { return is_const() ? P.cl[m].nu : If2<T, 2, 1>(P.g, P.nue[m], i, j, k); }
};

// Ax_q(viscous flux) etc: area * (-2 ν Σ)
template <typename T> struct VFlux {
    const TendP<T> &P;
    Grad<T> G;
    Visc<T> nu;
    __device__ VFlux(const TendP<T> &P_, int m) : P(P_), G{P_}, nu{P_, m} {}
    __device__ __forceinline__ T ux(int i, int j, int k) const { return (DYC * DZC(k)) * (-2 * (nu.ccc(i, j, k) * G.S11(i, j, k))); }
    __device__ __forceinline__ T uy(int i, int j, int k) const { return (DXF * DZC(k)) * (-2 * (nu.ffc(i, j, k) * G.S12(i, j, k))); }
    // VerticallyImplicitTimeDiscretization on a z-Bounded grid (abstract_scalar_diffusivity_closure.jl:270-312): away from the
    // boundary faces k = 1, Nz+1 only the part the tridiagonal solve does not contain stays explicit
    __device__ __forceinline__ bool elide(int k) const { return P.cl[nu.m].vi && P.g.topo[2] == BOUNDED && !((k == 1) | (k == P.g.N[2] + 1)); }
    __device__ __forceinline__ T uz(int i, int j, int k) const {
        if (elide(k)) return (DXF * DYC) * (-(nu.fcf(i, j, k) * G.dx_w(i, j, k)));
        return (DXF * DYC) * (-2 * (nu.fcf(i, j, k) * G.S13(i, j, k)));
    }
    __device__ __forceinline__ T vx(int i, int j, int k) const { return (DYF * DZC(k)) * (-2 * (nu.ffc(i, j, k) * G.S12(i, j, k))); }
    __device__ __forceinline__ T vy(int i, int j, int k) const { return (DXC * DZC(k)) * (-2 * (nu.ccc(i, j, k) * G.S22(i, j, k))); }
    __device__ __forceinline__ T vz(int i, int j, int k) const {
        if (elide(k)) return (DXC * DYF) * (-(nu.cff(i, j, k) * G.dy_w(i, j, k)));
        return (DXC * DYF) * (-2 * (nu.cff(i, j, k) * G.S23(i, j, k)));
    }
    __device__ __forceinline__ T wx(int i, int j, int k) const { return (DYC * DZF(k)) * (-2 * (nu.fcf(i, j, k) * G.S13(i, j, k))); }
    __device__ __forceinline__ T wy(int i, int j, int k) const { return (DXC * DZF(k)) * (-2 * (nu.cff(i, j, k) * G.S23(i, j, k))); }
    __device__ __forceinline__ T wz(int i, int j, int k) const {
        if (elide(k)) return T(0);
        return (DXC * DYC) * (-2 * (nu.ccc(i, j, k) * G.S33(i, j, k)));
    }
};

#define DELTA_(hi, lo, d) (P.g.topo[d] == FLAT ? T(0) : ((hi) - (lo)))

template <typename T> __device__ __forceinline__ T div_tau1(const TendP<T> &P, int m, int i, int j, int k) {
    VFlux<T> F(P, m);
    T Vi = P.g.rVc(k);
    return mul_rn(Vi, DELTA_(F.ux(i, j, k), F.ux(i - 1, j, k), 0) + DELTA_(F.uy(i, j + 1, k), F.uy(i, j, k), 1) +
                 DELTA_(F.uz(i, j, k + 1), F.uz(i, j, k), 2));
}
template <typename T> __device__ __forceinline__ T div_tau2(const TendP<T> &P, int m, int i, int j, int k) {
    VFlux<T> F(P, m);
    T Vi = P.g.rVc(k);
    return mul_rn(Vi, DELTA_(F.vx(i + 1, j, k), F.vx(i, j, k), 0) + DELTA_(F.vy(i, j, k), F.vy(i, j - 1, k), 1) +
                 DELTA_(F.vz(i, j, k + 1), F.vz(i, j, k), 2));
}
template <typename T> __device__ __forceinline__ T div_tau3(const TendP<T> &P, int m, int i, int j, int k) {
    VFlux<T> F(P, m);
    T Vi = P.g.rVf(k);
    return mul_rn(Vi, DELTA_(F.wx(i + 1, j, k), F.wx(i, j, k), 0) + DELTA_(F.wy(i, j + 1, k), F.wy(i, j, k), 1) +
                 DELTA_(F.wz(i, j, k), F.wz(i, j, k - 1), 2));
}

// diffusivity at fcc / cfc / ccf (D = 0,1,2)
template <typename T, int D> __device__ __forceinline__ T kap(const TendP<T> &P, int m, int t, int i, int j, int k) {
    const int kind = P.cl[m].kind;
    if (kind == CL_SCALAR) return P.cl[m].kappa[t];
    if (kind == CL_SMAG) return If1<T, D>(P.g, P.nue[m], i, j, k) / P.cl[m].Pr[t];
    return If1<T, D>(P.g, P.kappae[m][t], i, j, k);
}
template <typename T, int D> __device__ __forceinline__ T qflux(const TendP<T> &P, int m, int t, int i, int j, int k) {
    const Fld<T> &c = P.c[t];
    int a = i, b = j, cc = k;
    shift<D>(a, b, cc, -1);
    T rd = D == 0 ? P.g.rdx : D == 1 ? P.g.rdy : P.g.rdzF(k);
    T A = D == 0 ? DYC * DZC(k) : D == 1 ? DXC * DZC(k) : DXC * DYC;
    if (D == 2 && P.cl[m].vi && P.g.topo[2] == BOUNDED && !((k == 1) | (k == P.g.N[2] + 1))) return T(0);   // diffusive_flux_z(::VITD)
    T dc = (P.g.topo[D] == FLAT ? T(0) : c.ld(i, j, k) - c.ld(a, b, cc)) * rd;
    return A * (-kap<T, D>(P, m, t, i, j, k) * dc);
}
template <typename T> __device__ __forceinline__ T div_q(const TendP<T> &P, int m, int t, int i, int j, int k) {
    T Vi = P.g.rVc(k);
    return mul_rn(Vi, DELTA_((qflux<T, 0>(P, m, t, i + 1, j, k)), (qflux<T, 0>(P, m, t, i, j, k)), 0) +
                 DELTA_((qflux<T, 1>(P, m, t, i, j + 1, k)), (qflux<T, 1>(P, m, t, i, j, k)), 1) +
                 DELTA_((qflux<T, 2>(P, m, t, i, j, k + 1)), (qflux<T, 2>(P, m, t, i, j, k)), 2));
}

// buoyancy_perturbationᶜᶜᶜ
template <typename T> __device__ __forceinline__ T bpert(const TendP<T> &P, int i, int j, int k) {
    if (P.buoy == BUOY_TRACER) return P.c[P.ib].ld(i, j, k);
    if (P.buoy == BUOY_SEAWATER) return P.grav * (P.alpha * P.c[P.iT].ld(i, j, k) - P.beta * P.c[P.iS].ld(i, j, k));
    return 0;
}

// ---- pointwise tendencies (nonhydrostatic_tendency_kernel_functions.jl:71-302) -----------------------------------
// Every tendency = finish(advective flux divergence); the generic kernel evaluates each face flux through the
// difference operator as the reference does, the marching kernel (tendency_tiled.cuh) shares face fluxes between cells.
template <typename T, class S, bool FAST> __device__ __forceinline__ T Gu_adv(const TendP<T> &P, int i, int j, int k) {
    T Vi = P.g.rVc(k);
    return Vi * (DELTA_((mom_flux<T, S, FAST, 0, 0>(P, i, j, k)), (mom_flux<T, S, FAST, 0, 0>(P, i - 1, j, k)), 0) +
                 DELTA_((mom_flux<T, S, FAST, 1, 0>(P, i, j + 1, k)), (mom_flux<T, S, FAST, 1, 0>(P, i, j, k)), 1) +
                 DELTA_((mom_flux<T, S, FAST, 2, 0>(P, i, j, k + 1)), (mom_flux<T, S, FAST, 2, 0>(P, i, j, k)), 2));
}
template <typename T> __device__ __forceinline__ T Gu_finish(const TendP<T> &P, T adv, int i, int j, int k) {
    T r = -adv;
    if (P.has_cor) {  // coriolis_schemes.jl:67 : -ℑy(f) * ℑxyᶠᶜᵃ(Ay_q v) * Ay⁻¹ᶠᶜᶜ
        const bool fy = P.g.topo[1] == FLAT, fx = P.g.topo[0] == FLAT;
        T fbar = fy ? P.f : T(0.5) * (P.f + P.f);
        // products and sums rounded one by one, as the reference evaluates them (interp4_rn, common.cuh)
        auto Ayv = [&](int a, int b) { return mul_rn(DXC * DZC(k), P.v.ld(a, b, k)); };
        auto Ix = [&](int b) { return fx ? Ayv(i, b) : mul_rn(T(0.5), add_rn(Ayv(i - 1, b), Ayv(i, b))); };
        T I = fy ? Ix(j) : mul_rn(T(0.5), add_rn(Ix(j), Ix(j + 1)));
        r = sub_rn(r, mul_rn(mul_rn(-fbar, I), 1 / (DXF * DZC(k))));
    }
    // (-∂x pHY' rounded product by product, as the reference: every kernel form uses the same pinned expression)
    if (P.has_pHY) r = sub_rn(r, mul_rn(P.g.topo[0] == FLAT ? T(0) : P.pHY.ld(i, j, k) - P.pHY.ld(i - 1, j, k), P.g.rdx));
    if (P.ncl > 0) {
        T t = div_tau1(P, 0, i, j, k);
        for (int m = 1; m < P.ncl; m++) t = add_rn(t, div_tau1(P, m, i, j, k));
        r = r - t;
    }
    return r;
}
template <typename T, class S, bool FAST> __device__ __forceinline__ T Gv_adv(const TendP<T> &P, int i, int j, int k) {
    T Vi = P.g.rVc(k);
    return Vi * (DELTA_((mom_flux<T, S, FAST, 0, 1>(P, i + 1, j, k)), (mom_flux<T, S, FAST, 0, 1>(P, i, j, k)), 0) +
                 DELTA_((mom_flux<T, S, FAST, 1, 1>(P, i, j, k)), (mom_flux<T, S, FAST, 1, 1>(P, i, j - 1, k)), 1) +
                 DELTA_((mom_flux<T, S, FAST, 2, 1>(P, i, j, k + 1)), (mom_flux<T, S, FAST, 2, 1>(P, i, j, k)), 2));
}
template <typename T> __device__ __forceinline__ T Gv_finish(const TendP<T> &P, T adv, int i, int j, int k) {
    T r = -adv;
    if (P.has_cor) {  // coriolis_schemes.jl:68 : +ℑx(f) * ℑxyᶜᶠᵃ(Ax_q u) * Ax⁻¹ᶜᶠᶜ
        const bool fy = P.g.topo[1] == FLAT, fx = P.g.topo[0] == FLAT;
        T fbar = fx ? P.f : T(0.5) * (P.f + P.f);
        auto Axu = [&](int a, int b) { return mul_rn(DYC * DZC(k), P.u.ld(a, b, k)); };
        auto Ix = [&](int b) { return fx ? Axu(i, b) : mul_rn(T(0.5), add_rn(Axu(i, b), Axu(i + 1, b))); };
        T I = fy ? Ix(j) : mul_rn(T(0.5), add_rn(Ix(j - 1), Ix(j)));
        r = sub_rn(r, mul_rn(mul_rn(fbar, I), 1 / (DYF * DZC(k))));
    }
    if (P.has_pHY) r = sub_rn(r, mul_rn(P.g.topo[1] == FLAT ? T(0) : P.pHY.ld(i, j, k) - P.pHY.ld(i, j - 1, k), P.g.rdy));
    if (P.ncl > 0) {
        T t = div_tau2(P, 0, i, j, k);
        for (int m = 1; m < P.ncl; m++) t = add_rn(t, div_tau2(P, m, i, j, k));
        r = r - t;
    }
    return r;
}
template <typename T, class S, bool FAST> __device__ __forceinline__ T Gw_adv(const TendP<T> &P, int i, int j, int k) {
    T Vi = P.g.rVf(k);
    return Vi * (DELTA_((mom_flux<T, S, FAST, 0, 2>(P, i + 1, j, k)), (mom_flux<T, S, FAST, 0, 2>(P, i, j, k)), 0) +
                 DELTA_((mom_flux<T, S, FAST, 1, 2>(P, i, j + 1, k)), (mom_flux<T, S, FAST, 1, 2>(P, i, j, k)), 1) +
                 DELTA_((mom_flux<T, S, FAST, 2, 2>(P, i, j, k)), (mom_flux<T, S, FAST, 2, 2>(P, i, j, k - 1)), 2));
}
template <typename T> __device__ __forceinline__ T Gw_finish(const TendP<T> &P, T adv, int i, int j, int k) {
    T r = -adv;
    if (!P.has_pHY && P.buoy != BUOY_NONE) {  // maybe_z_dot_g_bᶜᶜᶠ = ℑzᵃᵃᶠ(b)
        r = r + (P.g.topo[2] == FLAT ? bpert(P, i, j, k) : T(0.5) * (bpert(P, i, j, k - 1) + bpert(P, i, j, k)));
    }
    if (P.ncl > 0) {
        T t = div_tau3(P, 0, i, j, k);
        for (int m = 1; m < P.ncl; m++) t = add_rn(t, div_tau3(P, m, i, j, k));
        r = r - t;
    }
    return r;
}
template <typename T, class S, bool FAST> __device__ __forceinline__ T Gc_adv(const TendP<T> &P, int t, int i, int j, int k) {
    const Fld<T> &c = P.c[t];
    T Vi = P.g.rVc(k);
    return Vi * (DELTA_((tracer_flux<T, S, FAST, 0>(P, c, i + 1, j, k)), (tracer_flux<T, S, FAST, 0>(P, c, i, j, k)), 0) +
                 DELTA_((tracer_flux<T, S, FAST, 1>(P, c, i, j + 1, k)), (tracer_flux<T, S, FAST, 1>(P, c, i, j, k)), 1) +
                 DELTA_((tracer_flux<T, S, FAST, 2>(P, c, i, j, k + 1)), (tracer_flux<T, S, FAST, 2>(P, c, i, j, k)), 2));
}
template <typename T> __device__ __forceinline__ T Gc_finish(const TendP<T> &P, T adv, int t, int i, int j, int k) {
    T r = -adv;
    if (P.ncl > 0) {
        T q = div_q(P, 0, t, i, j, k);
        for (int m = 1; m < P.ncl; m++) q = add_rn(q, div_q(P, m, t, i, j, k));
        r = r - q;
    }
    return r;
}

// Generic fused tendency kernel.  Thread (x fastest) <-> cell; blockIdx.z selects the tendency so that the four
// instruction streams do not share registers (the reference's 3+n launches become one launch).
template <typename T, class S, bool FAST>
__global__ void __launch_bounds__(128) tendency_generic_kernel(const __grid_constant__ TendP<T> P, int i0, int i1) {
    int i, j, k;
    if (!cell_from_block(i1 - i0 + 1, P.g.N[1], i, j, k)) return;
    i += i0 - 1;
    const int which = blockIdx.y;
    if (which == 0) P.Gu(i, j, k) = Gu_finish<T>(P, Gu_adv<T, S, FAST>(P, i, j, k), i, j, k);
    else if (which == 1) P.Gv(i, j, k) = Gv_finish<T>(P, Gv_adv<T, S, FAST>(P, i, j, k), i, j, k);
    else if (which == 2) P.Gw(i, j, k) = Gw_finish<T>(P, Gw_adv<T, S, FAST>(P, i, j, k), i, j, k);
    else P.Gc[which - 3](i, j, k) = Gc_finish<T>(P, Gc_adv<T, S, FAST>(P, which - 3, i, j, k), which - 3, i, j, k);
}

#undef DXF
#undef DXC
#undef DYF
#undef DYC
#undef DZF
#undef DZC

}  // namespace ob
