// tendency_fast.cuh -- interior fast path of the flux-sharing marching kernel for WENO schemes.
//
// Same decomposition and the same arithmetic as march_body (tendency_tiled.cuh), specialised for CTAs whose
// advective stencils never touch a Bounded boundary: x and y Periodic (all parent arrays then share the same row and
// plane strides), z Periodic or the CTA's k-chunk away from the walls.  There the Bounded fallback chain
// (topologically_conditional_interpolation.jl:43-128) always selects the full scheme, so the flux reduces to
//     momentum:  ũ = Centered(2N-2) interpolation of (A U) ; q̂ = WENO(2N-1) of q with bias sign(ũ) ; F = ũ q̂
//     tracer:    F = A u ĉ, bias sign(u)                       (upwind_biased_advective_fluxes.jl:23-121)
// and everything is addressed from ONE base pointer per field with compile-time x offsets and two runtime strides.
// The generic per-point functions (runtime topology, 64-bit index arithmetic per load) cost ~3 integer/constant
// instructions per FP64 instruction; this path removes most of them (profiles/ r1 notes).
#pragma once
#include "tendency.cuh"

namespace ob {

// LIT: coefficients of Centered(2 NC) = the advecting scheme of WENO(NC + 1) as literals (coef_literals.h)
template <typename T, int NC, bool LIT = false>
__device__ __forceinline__ T centered_vals(const T (&v)[2 * NC]) {
    const auto &tab = Tab<T>::get();
    auto cc = [&](int m) -> T { if constexpr (LIT) return CoefLit<T, NC + 1>::cen(m); else return tab.cen_coeff[NC][m]; };
    T acc = cc(0) * v[0];
#pragma unroll
    for (int m = 1; m < 2 * NC; m++) acc = fma_(cc(m), v[m], acc);
    return acc;
}

// WENO(2N-1) from the 2N values s[m] = ψ[face - N + m]; LeftBias uses s[0 .. 2N-2], RightBias the mirror image
template <typename T, int N, bool FAST, bool LIT = false>
__device__ __forceinline__ T weno_sel(const T (&s)[2 * N], bool left) {
    T v[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) v[m] = left ? s[m] : s[2 * N - 1 - m];
    return weno_from_values<T, N, FAST, LIT>(v);
}

// Geometry of the fast path.  STR = stretched z: metrics are read per level from the host-built arrays; otherwise every
// metric is a kernel constant (no pointer tests, no loads).
template <typename T, bool STR>
struct FastGeom {
    int sy, sz;           // row / plane strides in elements (identical for every field on this path; parents < 2^31 elements)
    T dx, dy, rdx, rdy;
    T dz, rdz, rvol;
    const T *dzc, *dzf, *rdzc, *rdzf, *rvc, *rvf;  // pre-offset, logical k
    __device__ __forceinline__ T dzC(int k) const { if constexpr (STR) return __ldg(dzc + k); else return dz; }
    __device__ __forceinline__ T dzF(int k) const { if constexpr (STR) return __ldg(dzf + k); else return dz; }
    __device__ __forceinline__ T rdzC(int k) const { if constexpr (STR) return __ldg(rdzc + k); else return rdz; }
    __device__ __forceinline__ T rdzF(int k) const { if constexpr (STR) return __ldg(rdzf + k); else return rdz; }
    __device__ __forceinline__ T rVc(int k) const { if constexpr (STR) return __ldg(rvc + k); else return rvol; }
    __device__ __forceinline__ T rVf(int k) const { if constexpr (STR) return __ldg(rvf + k); else return rvol; }
    __device__ __forceinline__ void init(const GridD<T> &gg, int sy_, long sz_) {
        sy = sy_; sz = (int)sz_; dx = gg.dx; dy = gg.dy; rdx = gg.rdx; rdy = gg.rdy; dz = gg.dz; rdz = gg.rdz; rvol = gg.rvol;
        dzc = gg.dzc; dzf = gg.dzf; rdzc = gg.rdzc; rdzf = gg.rdzf; rvc = gg.rvc; rvf = gg.rvf;
    }
};

template <int DIR, typename T, bool STR>
__device__ __forceinline__ int stride_of(const FastGeom<T, STR> &g) { return DIR == 0 ? 1 : DIR == 1 ? g.sy : g.sz; }

// q[m] = p[(lo + m) * stride<DIR>], m = 0 .. CNT-1
template <int DIR, int CNT, typename T, bool STR>
__device__ __forceinline__ void load_line(const T *__restrict__ p, const FastGeom<T, STR> &g, int lo, T (&out)[CNT]) {
    const int st = stride_of<DIR>(g);
#pragma unroll
    for (int m = 0; m < CNT; m++) out[m] = __ldg(p + (lo + m) * st);
}

// Advective flux in direction ADV of tendency WHICH for the thread whose own point is (i, j, kp) -- kp = k for the
// x/y fluxes, k+1 for the upper z flux.  pq / pa point at (i, j, kp) of the advected field and of the advecting
// velocity component ADV.
// The flux from already-loaded stencil values: s[m] = q at offsets -N .. N-1 along ADV; a[] = the advecting component
// at offsets -(N-1) .. N-2 along axis WHICH (momentum) or its single value at the face (tracer).
template <typename T, int N, bool FAST, int WHICH, int ADV, bool STR, bool LIT = false>
__device__ __forceinline__ T flux_from_values(const T (&s)[2 * N], T (&a)[WHICH == 3 ? 1 : 2 * (N - 1)], const FastGeom<T, STR> &g, int kp) {
    if constexpr (WHICH == 3) {
        const T A = ADV == 0 ? g.dy * g.dzC(kp) : ADV == 1 ? g.dx * g.dzC(kp) : g.dx * g.dy;
        const T ut = a[0];
        const T cr = weno_sel<T, N, FAST, LIT>(s, ut > 0);
        return A * ut * cr;
    } else {
        constexpr int NC = N - 1;
#pragma unroll
        for (int m = 0; m < 2 * NC; m++) {
            // Ax_qᶠᶜᶜ = Δy Δzᶜ(k') u, Ay_qᶜᶠᶜ = Δx Δzᶜ(k') v, Az_qᶜᶜᶠ = Δx Δy w, k' the level of the stencil point
            const int kq = WHICH == 2 ? kp + m - NC : kp;
            const T A = ADV == 0 ? g.dy * g.dzC(kq) : ADV == 1 ? g.dx * g.dzC(kq) : g.dx * g.dy;
            a[m] = A * a[m];
        }
        const T ut = centered_vals<T, NC, LIT>(a);
        const T qr = weno_sel<T, N, FAST, LIT>(s, ut > 0);
        return ut * qr;
    }
}

template <typename T, int N, bool FAST, int WHICH, int ADV, bool STR>
__device__ __forceinline__ T fast_flux(const T *__restrict__ pq, const T *__restrict__ pa, const FastGeom<T, STR> &g, int kp) {
    // (same arithmetic as flux_from_values, written out: this form measured 1.5 % faster in the LDG kernel)
    T s[2 * N];
    load_line<ADV, 2 * N>(pq, g, -N, s);
    if constexpr (WHICH == 3) {
        const T A = ADV == 0 ? g.dy * g.dzC(kp) : ADV == 1 ? g.dx * g.dzC(kp) : g.dx * g.dy;
        const T ut = __ldg(pa);
        const T cr = weno_sel<T, N, FAST>(s, ut > 0);
        return A * ut * cr;
    } else {
        constexpr int NC = N - 1;
        T a[2 * NC];
        load_line<WHICH, 2 * NC>(pa, g, -NC, a);
#pragma unroll
        for (int m = 0; m < 2 * NC; m++) {
            const int kq = WHICH == 2 ? kp + m - NC : kp;
            const T A = ADV == 0 ? g.dy * g.dzC(kq) : ADV == 1 ? g.dx * g.dzC(kq) : g.dx * g.dy;
            a[m] = A * a[m];
        }
        const T ut = centered_vals<T, NC>(a);
        const T qr = weno_sel<T, N, FAST>(s, ut > 0);
        return ut * qr;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Non-advective terms on the fast path: the arithmetic of Gu/Gv/Gw/Gc_finish (tendency.cuh) restated with offsets
// from the thread's own point (a, b, c are compile-time after inlining) so that every load is base + constant.
// Valid where x, y, z are not Flat and every field shares the strides (sy, sz).
// ------------------------------------------------------------------------------------------------------------------
template <typename T, bool STR>
struct FastTerms {
    const TendP<T> &P;
    const FastGeom<T, STR> &G;
    const T *u, *v, *w;   // at the thread's (i, j, k)
    int eo;               // element offset of (i, j, k) relative to logical (0, 0, 0): i + j*sy + k*sz
    int k;
    __device__ __forceinline__ T ld(const T *p, int a, int b, int c) const { return __ldg(p + (a + b * G.sy + c * G.sz)); }
    __device__ __forceinline__ const T *at(const Fld<T> &f) const { return f.p + f.off + eo; }
    __device__ __forceinline__ T dzC(int c) const { return G.dzC(k + c); }
    __device__ __forceinline__ T dzF(int c) const { return G.dzF(k + c); }
    // velocity gradients (velocity_tracer_gradients.jl:6-19)
    __device__ __forceinline__ T dx_u(int a, int b, int c) const { return (ld(u, a + 1, b, c) - ld(u, a, b, c)) * G.rdx; }
    __device__ __forceinline__ T dy_v(int a, int b, int c) const { return (ld(v, a, b + 1, c) - ld(v, a, b, c)) * G.rdy; }
    __device__ __forceinline__ T dz_w(int a, int b, int c) const { return (ld(w, a, b, c + 1) - ld(w, a, b, c)) * G.rdzC(k + c); }
    __device__ __forceinline__ T dx_v(int a, int b, int c) const { return (ld(v, a, b, c) - ld(v, a - 1, b, c)) * G.rdx; }
    __device__ __forceinline__ T dy_u(int a, int b, int c) const { return (ld(u, a, b, c) - ld(u, a, b - 1, c)) * G.rdy; }
    __device__ __forceinline__ T dx_w(int a, int b, int c) const { return (ld(w, a, b, c) - ld(w, a - 1, b, c)) * G.rdx; }
    __device__ __forceinline__ T dz_u(int a, int b, int c) const { return (ld(u, a, b, c) - ld(u, a, b, c - 1)) * G.rdzF(k + c); }
    __device__ __forceinline__ T dy_w(int a, int b, int c) const { return (ld(w, a, b, c) - ld(w, a, b - 1, c)) * G.rdy; }
    __device__ __forceinline__ T dz_v(int a, int b, int c) const { return (ld(v, a, b, c) - ld(v, a, b, c - 1)) * G.rdzF(k + c); }
    __device__ __forceinline__ T S12(int a, int b, int c) const { return T(0.5) * add_rn(dy_u(a, b, c), dx_v(a, b, c)); }
    __device__ __forceinline__ T S13(int a, int b, int c) const { return T(0.5) * add_rn(dz_u(a, b, c), dx_w(a, b, c)); }
    __device__ __forceinline__ T S23(int a, int b, int c) const { return T(0.5) * add_rn(dz_v(a, b, c), dy_w(a, b, c)); }
    // ℑ of a ccc array to faces (interpolation_operators.jl:8-71): D = direction of the -1 shift
    __device__ __forceinline__ T If1(const T *f, int D, int a, int b, int c) const {
        return T(0.5) * (ld(f, a - (D == 0), b - (D == 1), c - (D == 2)) + ld(f, a, b, c));
    }
    __device__ __forceinline__ T If2(const T *f, int D2, int D1, int a, int b, int c) const {
        return T(0.5) * (If1(f, D1, a - (D2 == 0), b - (D2 == 1), c - (D2 == 2)) + If1(f, D1, a, b, c));
    }
    // viscosity of closure m at ccc / ffc / fcf / cff (abstract_scalar_diffusivity_closure.jl:330-351)
    __device__ __forceinline__ T nu_ccc(int m, const T *ne, int a, int b, int c) const { return ne ? ld(ne, a, b, c) : P.cl[m].nu; }
    __device__ __forceinline__ T nu_ffc(int m, const T *ne, int a, int b, int c) const { return ne ? If2(ne, 1, 0, a, b, c) : P.cl[m].nu; }
    __device__ __forceinline__ T nu_fcf(int m, const T *ne, int a, int b, int c) const { return ne ? If2(ne, 2, 0, a, b, c) : P.cl[m].nu; }
    __device__ __forceinline__ T nu_cff(int m, const T *ne, int a, int b, int c) const { return ne ? If2(ne, 2, 1, a, b, c) : P.cl[m].nu; }
    // Ax_q(viscous flux) = area * (-2 ν Σ) (closure_kernel_operators.jl:20-40)
    __device__ __forceinline__ T ux(int m, const T *ne, int a, int b, int c) const { return (G.dy * dzC(c)) * (-2 * (nu_ccc(m, ne, a, b, c) * dx_u(a, b, c))); }
    __device__ __forceinline__ T uy(int m, const T *ne, int a, int b, int c) const { return (G.dx * dzC(c)) * (-2 * (nu_ffc(m, ne, a, b, c) * S12(a, b, c))); }
    __device__ __forceinline__ T uz(int m, const T *ne, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_fcf(m, ne, a, b, c) * S13(a, b, c))); }
    __device__ __forceinline__ T vx(int m, const T *ne, int a, int b, int c) const { return (G.dy * dzC(c)) * (-2 * (nu_ffc(m, ne, a, b, c) * S12(a, b, c))); }
    __device__ __forceinline__ T vy(int m, const T *ne, int a, int b, int c) const { return (G.dx * dzC(c)) * (-2 * (nu_ccc(m, ne, a, b, c) * dy_v(a, b, c))); }
    __device__ __forceinline__ T vz(int m, const T *ne, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_cff(m, ne, a, b, c) * S23(a, b, c))); }
    __device__ __forceinline__ T wx(int m, const T *ne, int a, int b, int c) const { return (G.dy * dzF(c)) * (-2 * (nu_fcf(m, ne, a, b, c) * S13(a, b, c))); }
    __device__ __forceinline__ T wy(int m, const T *ne, int a, int b, int c) const { return (G.dx * dzF(c)) * (-2 * (nu_cff(m, ne, a, b, c) * S23(a, b, c))); }
    __device__ __forceinline__ T wz(int m, const T *ne, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_ccc(m, ne, a, b, c) * dz_w(a, b, c))); }
    __device__ __forceinline__ const T *nue_ptr(int m) const { return P.cl[m].kind == CL_SCALAR ? nullptr : at(P.nue[m]); }
    template <int WHICH> __device__ __forceinline__ T div_tau(int m) const {
        const T *ne = nue_ptr(m);
        if constexpr (WHICH == 0)
            return mul_rn(G.rVc(k), (ux(m, ne, 0, 0, 0) - ux(m, ne, -1, 0, 0)) + (uy(m, ne, 0, 1, 0) - uy(m, ne, 0, 0, 0)) + (uz(m, ne, 0, 0, 1) - uz(m, ne, 0, 0, 0)));
        else if constexpr (WHICH == 1)
            return mul_rn(G.rVc(k), (vx(m, ne, 1, 0, 0) - vx(m, ne, 0, 0, 0)) + (vy(m, ne, 0, 0, 0) - vy(m, ne, 0, -1, 0)) + (vz(m, ne, 0, 0, 1) - vz(m, ne, 0, 0, 0)));
        else
            return mul_rn(G.rVf(k), (wx(m, ne, 1, 0, 0) - wx(m, ne, 0, 0, 0)) + (wy(m, ne, 0, 1, 0) - wy(m, ne, 0, 0, 0)) + (wz(m, ne, 0, 0, 0) - wz(m, ne, 0, 0, -1)));
    }
    // diffusive flux of tracer t along D at the face (a, b, c) (abstract_scalar_diffusivity_closure.jl:260-262)
    __device__ __forceinline__ T qflux(int m, int t, const T *cp, const T *kf, int D, int a, int b, int c) const {
        const int kind = P.cl[m].kind;
        const T kap = kind == CL_SCALAR ? P.cl[m].kappa[t] : kind == CL_SMAG ? If1(kf, D, a, b, c) / P.cl[m].Pr[t] : If1(kf, D, a, b, c);
        const T rd = D == 0 ? G.rdx : D == 1 ? G.rdy : G.rdzF(k + c);
        const T A = D == 0 ? G.dy * dzC(c) : D == 1 ? G.dx * dzC(c) : G.dx * G.dy;
        const T dc = (ld(cp, a, b, c) - ld(cp, a - (D == 0), b - (D == 1), c - (D == 2))) * rd;
        return A * (-kap * dc);
    }
    __device__ __forceinline__ T div_q(int m, int t, const T *cp) const {
        const int kind = P.cl[m].kind;
        const T *kf = kind == CL_SCALAR ? nullptr : kind == CL_SMAG ? at(P.nue[m]) : at(P.kappae[m][t]);
        return mul_rn(G.rVc(k), (qflux(m, t, cp, kf, 0, 1, 0, 0) - qflux(m, t, cp, kf, 0, 0, 0, 0)) + (qflux(m, t, cp, kf, 1, 0, 1, 0) - qflux(m, t, cp, kf, 1, 0, 0, 0)) +
                                    (qflux(m, t, cp, kf, 2, 0, 0, 1) - qflux(m, t, cp, kf, 2, 0, 0, 0)));
    }
    __device__ __forceinline__ T bpert(int c) const {
        if (P.buoy == BUOY_TRACER) return ld(at(P.c[P.ib]), 0, 0, c);
        if (P.buoy == BUOY_SEAWATER) return P.grav * (P.alpha * ld(at(P.c[P.iT]), 0, 0, c) - P.beta * ld(at(P.c[P.iS]), 0, 0, c));
        return 0;
    }
    // closure flux of tendency WHICH through the face this thread OWNS in direction D (same ownership as the advective
    // fluxes: centre-type directions own the face one index below; D == 2 is the UPPER face, evaluated from level k)
    template <int WHICH, int D> __device__ __forceinline__ T own_closure_flux(int m, int t, const T *cp) const {
        if constexpr (WHICH == 3) {
            const int kind = P.cl[m].kind;
            const T *kf = kind == CL_SCALAR ? nullptr : kind == CL_SMAG ? at(P.nue[m]) : at(P.kappae[m][t]);
            return qflux(m, t, cp, kf, D, 0, 0, D == 2 ? 1 : 0);
        } else {
            const T *ne = nue_ptr(m);
            if constexpr (WHICH == 0) return D == 0 ? ux(m, ne, -1, 0, 0) : D == 1 ? uy(m, ne, 0, 0, 0) : uz(m, ne, 0, 0, 1);
            else if constexpr (WHICH == 1) return D == 0 ? vx(m, ne, 0, 0, 0) : D == 1 ? vy(m, ne, 0, -1, 0) : vz(m, ne, 0, 0, 1);
            else return D == 0 ? wx(m, ne, 0, 0, 0) : D == 1 ? wy(m, ne, 0, 0, 0) : wz(m, ne, 0, 0, 0);
        }
    }
    // lower z face at the first level of a chunk
    template <int WHICH> __device__ __forceinline__ T first_lower_closure_flux(int m, int t, const T *cp) const {
        if constexpr (WHICH == 3) {
            const int kind = P.cl[m].kind;
            const T *kf = kind == CL_SCALAR ? nullptr : kind == CL_SMAG ? at(P.nue[m]) : at(P.kappae[m][t]);
            return qflux(m, t, cp, kf, 2, 0, 0, 0);
        } else {
            const T *ne = nue_ptr(m);
            if constexpr (WHICH == 0) return uz(m, ne, 0, 0, 0);
            else if constexpr (WHICH == 1) return vz(m, ne, 0, 0, 0);
            else return wz(m, ne, 0, 0, -1);
        }
    }
    // the tendency assemblers (nonhydrostatic_tendency_kernel_functions.jl:71-302), same term order as G*_finish;
    // SHARED: the closure term (sum over closures of V⁻¹ Σ δ(flux)) was formed from face fluxes shared between cells
    template <int WHICH, bool SHARED = false> __device__ __forceinline__ T finish(T adv, int t, const T *cp, T closure_term = T(0)) const {
        T r = -adv;
        if constexpr (WHICH == 0) {
            if (P.has_cor) {
                const T fbar = T(0.5) * (P.f + P.f);
                const T A = G.dx * dzC(0);
                const T I = interp4_rn(A, ld(v, -1, 0, 0), ld(v, 0, 0, 0), ld(v, -1, 1, 0), ld(v, 0, 1, 0));
                r = sub_rn(r, mul_rn(mul_rn(-fbar, I), 1 / (G.dx * dzC(0))));
            }
            if (P.has_pHY) { const T *ph = at(P.pHY); r = sub_rn(r, mul_rn(ld(ph, 0, 0, 0) - ld(ph, -1, 0, 0), G.rdx)); }
        } else if constexpr (WHICH == 1) {
            if (P.has_cor) {
                const T fbar = T(0.5) * (P.f + P.f);
                const T A = G.dy * dzC(0);
                const T I = interp4_rn(A, ld(u, 0, -1, 0), ld(u, 1, -1, 0), ld(u, 0, 0, 0), ld(u, 1, 0, 0));
                r = sub_rn(r, mul_rn(mul_rn(fbar, I), 1 / (G.dy * dzC(0))));
            }
            if (P.has_pHY) { const T *ph = at(P.pHY); r = sub_rn(r, mul_rn(ld(ph, 0, 0, 0) - ld(ph, 0, -1, 0), G.rdy)); }
        } else if constexpr (WHICH == 2) {
            if (!P.has_pHY && P.buoy != BUOY_NONE) r = r + T(0.5) * (bpert(-1) + bpert(0));
        }
        if (P.ncl > 0) {
            if constexpr (SHARED) {
                r = r - closure_term;
            } else if constexpr (WHICH == 3) {
                T q = div_q(0, t, cp);
                for (int m = 1; m < P.ncl; m++) q = add_rn(q, div_q(m, t, cp));
                r = r - q;
            } else {
                T tt = div_tau<WHICH>(0);
                for (int m = 1; m < P.ncl; m++) tt = add_rn(tt, div_tau<WHICH>(m));
                r = r - tt;
            }
        }
        return r;
    }
};

#define OB_SHARED_CL 2   // closures whose face fluxes are shared between cells on the fast path (more: pointwise)
template <typename T, int N, bool FAST, int WHICH, int TY, int KC, bool STR>
__device__ __forceinline__ void march_fast_body(const TendP<T> &P, int t, int i, int j, int k0, int k1, T (*sy_buf)[TY][32],
                                                T (*sv_buf)[OB_SHARED_CL][TY][32]) {
    const GridD<T> &gg = P.g;
    const int Nx = gg.N[0], Ny = gg.N[1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool do_out = (tx < 31) && (ty < TY - 1) && (i <= Nx) && (j <= Ny);
    const bool full_row = ty < TY - 1;  // warp-uniform: the overlap row only supplies its y-direction flux
    const Fld<T> &qf = WHICH == 0 ? P.u : WHICH == 1 ? P.v : WHICH == 2 ? P.w : P.c[t];
    const Fld<T> &G = WHICH == 0 ? P.Gu : WHICH == 1 ? P.Gv : WHICH == 2 ? P.Gw : P.Gc[t];
    FastGeom<T, STR> g;
    g.init(gg, P.u.sy, P.u.sz);
    // Lanes outside the flux region (i > Nx+1, j > Ny+1) are clamped onto its edge: they compute valid-but-unused
    // fluxes, which keeps each level ONE branch-free block in which the three independent WENO chains interleave.
    const int ii = min(i, Nx + 1), jj = min(j, Ny + 1);
    const int base = ii + jj * g.sy + k0 * g.sz;  // every field has the same offsets on this path
    const T *pq = qf.p + qf.off + base;
    const T *pu = P.u.p + P.u.off + base, *pv = P.v.p + P.v.off + base, *pw = P.w.p + P.w.off + base;
    // closure face fluxes are shared exactly like the advective ones when there are 1..OB_SHARED_CL closures
    const int ncl = P.ncl;
    const bool share_cl = ncl >= 1 && ncl <= OB_SHARED_CL;   // launch-uniform
    T lower = full_row ? fast_flux<T, N, FAST, WHICH, 2, STR>(pq, pw, g, k0) : T(0);
    T lower_c[OB_SHARED_CL] = {T(0), T(0)};
    if (share_cl && full_row) {
        FastTerms<T, STR> F0{P, g, pu, pv, pw, base, k0};
#pragma unroll
        for (int m = 0; m < OB_SHARED_CL; m++) if (m < ncl) lower_c[m] = F0.template first_lower_closure_flux<WHICH>(m, t, pq);
    }
    for (int k = k0; k <= k1; k++) {
        T fx = T(0), upper = T(0);
        const T fy = fast_flux<T, N, FAST, WHICH, 1, STR>(pq, pv, g, k);
        if (full_row) {
            fx = fast_flux<T, N, FAST, WHICH, 0, STR>(pq, pu, g, k);
            upper = fast_flux<T, N, FAST, WHICH, 2, STR>(pq + g.sz, pw + g.sz, g, k + 1);
        }
        const int eo = ii + jj * g.sy + k * g.sz;
        FastTerms<T, STR> F{P, g, pu, pv, pw, eo, k};
        T cx[OB_SHARED_CL] = {T(0), T(0)}, cy[OB_SHARED_CL] = {T(0), T(0)}, cup[OB_SHARED_CL] = {T(0), T(0)};
        if (share_cl) {
#pragma unroll
            for (int m = 0; m < OB_SHARED_CL; m++)
                if (m < ncl) {
                    cy[m] = F.template own_closure_flux<WHICH, 1>(m, t, pq);
                    if (full_row) {
                        cx[m] = F.template own_closure_flux<WHICH, 0>(m, t, pq);
                        cup[m] = F.template own_closure_flux<WHICH, 2>(m, t, pq);
                    }
                }
        }
        const T fx1 = __shfl_down_sync(0xffffffffu, fx, 1);
        T cx1[OB_SHARED_CL];
#pragma unroll
        for (int m = 0; m < OB_SHARED_CL; m++) cx1[m] = share_cl ? __shfl_down_sync(0xffffffffu, cx[m], 1) : T(0);
        const int buf = k & 1;
        sy_buf[buf][ty][tx] = fy;
        if (share_cl) {
#pragma unroll
            for (int m = 0; m < OB_SHARED_CL; m++) if (m < ncl) sv_buf[buf][m][ty][tx] = cy[m];
        }
        __syncthreads();
        if (do_out) {
            const T fy1 = sy_buf[buf][ty + 1][tx];
            const T Vi = WHICH == 2 ? g.rVf(k) : g.rVc(k);
            const T adv = Vi * ((fx1 - fx) + (fy1 - fy) + (upper - lower));
            T r;
            if (share_cl) {
                T term = T(0);
#pragma unroll
                for (int m = 0; m < OB_SHARED_CL; m++)
                    if (m < ncl) {
                        const T cy1 = sv_buf[buf][m][ty + 1][tx];
                        // every closure's flux divergence rounded on its own, then summed (the reference adds ∇·τ of the closures
                        // of a tuple one by one): the same pinned form in every kernel
                        const T d = mul_rn(Vi, (cx1[m] - cx[m]) + (cy1 - cy[m]) + (cup[m] - lower_c[m]));
                        term = m == 0 ? d : add_rn(term, d);
                    }
                r = F.template finish<WHICH, true>(adv, t, pq, term);
            } else {
                r = F.template finish<WHICH, false>(adv, t, pq);
            }
            G.p[G.off + eo] = r;
        }
        lower = upper;
#pragma unroll
        for (int m = 0; m < OB_SHARED_CL; m++) lower_c[m] = cup[m];
        pq += g.sz; pu += g.sz; pv += g.sz; pw += g.sz;
    }
}

// CTA-uniform predicate: every advective stencil of this k-chunk takes the full scheme
template <typename T, int N>
__device__ __forceinline__ bool fast_path_ok(const TendP<T> &P, int k0, int k1) {
    const GridD<T> &g = P.g;
    if (g.topo[0] != PERIODIC || g.topo[1] != PERIODIC) return false;
    // vertically-implicit closures take the generic flux functions: selecting the elided forms at run time inside the fast
    // body costs every model ~2 % of the kernel (predicated dual paths), a template parameter would double its code
    for (int m = 0; m < P.ncl; m++) if (P.cl[m].vi) return false;
    if (P.u.sz * (long)(g.N[2] + 2 * g.H[2] + 1) >= 2147483647L) return false;  // 32-bit element offsets
    if (g.topo[2] == PERIODIC) return true;
    if (g.topo[2] == FLAT) return false;
    // Bounded z: face-type flux indices k0 .. k1+1 and centre-type k0-1 .. k1 must satisfy outside_*_halo for buffer N
    return (k0 >= N + 1) && (k1 + 1 <= g.N[2] + 1 - N);
}

}  // namespace ob
