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

template <typename T, int NC>
__device__ __forceinline__ T centered_vals(const T (&v)[2 * NC]) {
    const auto &tab = Tab<T>::get();
    T acc = tab.cen_coeff[NC][0] * v[0];
#pragma unroll
    for (int m = 1; m < 2 * NC; m++) acc = fma_(tab.cen_coeff[NC][m], v[m], acc);
    return acc;
}

// WENO(2N-1) from the 2N values s[m] = ψ[face - N + m]; LeftBias uses s[0 .. 2N-2], RightBias the mirror image
template <typename T, int N, bool FAST>
__device__ __forceinline__ T weno_sel(const T (&s)[2 * N], bool left) {
    T v[2 * N - 1];
#pragma unroll
    for (int m = 0; m < 2 * N - 1; m++) v[m] = left ? s[m] : s[2 * N - 1 - m];
    return weno_from_values<T, N, FAST>(v);
}

template <typename T>
struct FastGeom {
    long sy, sz;          // row / plane strides in elements (identical for every field on this path)
    T dx, dy, dz;
    const T *dzc, *dzf;   // stretched z (pre-offset, logical k) or nullptr
    __device__ __forceinline__ T dzC(int k) const { return dzc ? __ldg(dzc + k) : dz; }
    __device__ __forceinline__ T dzF(int k) const { return dzf ? __ldg(dzf + k) : dz; }
};

template <int DIR, typename T>
__device__ __forceinline__ long stride_of(const FastGeom<T> &g) { return DIR == 0 ? 1L : DIR == 1 ? g.sy : g.sz; }

// q[m] = p[(lo + m) * stride<DIR>], m = 0 .. CNT-1
template <int DIR, int CNT, typename T>
__device__ __forceinline__ void load_line(const T *__restrict__ p, const FastGeom<T> &g, int lo, T (&out)[CNT]) {
    const long st = stride_of<DIR>(g);
#pragma unroll
    for (int m = 0; m < CNT; m++) out[m] = __ldg(p + (long)(lo + m) * st);
}

// Advective flux in direction ADV of tendency WHICH for the thread whose own point is (i, j, kp) -- kp = k for the
// x/y fluxes, k+1 for the upper z flux.  pq / pa point at (i, j, kp) of the advected field and of the advecting
// velocity component ADV.
template <typename T, int N, bool FAST, int WHICH, int ADV>
__device__ __forceinline__ T fast_flux(const T *__restrict__ pq, const T *__restrict__ pa, const FastGeom<T> &g, int kp) {
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
            // Ax_qᶠᶜᶜ = Δy Δzᶜ(k') u, Ay_qᶜᶠᶜ = Δx Δzᶜ(k') v, Az_qᶜᶜᶠ = Δx Δy w, k' the level of the stencil point
            const int kq = WHICH == 2 ? kp + m - NC : kp;
            const T A = ADV == 0 ? g.dy * g.dzC(kq) : ADV == 1 ? g.dx * g.dzC(kq) : g.dx * g.dy;
            a[m] = A * a[m];
        }
        const T ut = centered_vals<T, NC>(a);
        const T qr = weno_sel<T, N, FAST>(s, ut > 0);
        return ut * qr;
    }
}

template <typename T, int N, bool FAST, int WHICH, int TY, int KC>
__device__ __forceinline__ void march_fast_body(const TendP<T> &P, int t, int i, int j, int k0, int k1, T (*sy_buf)[TY][32]) {
    const GridD<T> &gg = P.g;
    const int Nx = gg.N[0], Ny = gg.N[1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool do_x = (ty < TY - 1) && (j <= Ny) && (i <= Nx + 1);
    const bool do_y = (tx < 31) && (i <= Nx) && (j <= Ny + 1);
    const bool do_out = (tx < 31) && (ty < TY - 1) && (i <= Nx) && (j <= Ny);
    const bool in_range = (i <= Nx + 1) && (j <= Ny + 1);
    const Fld<T> &qf = WHICH == 0 ? P.u : WHICH == 1 ? P.v : WHICH == 2 ? P.w : P.c[t];
    const Fld<T> &G = WHICH == 0 ? P.Gu : WHICH == 1 ? P.Gv : WHICH == 2 ? P.Gw : P.Gc[t];
    FastGeom<T> g;
    g.sy = P.u.sy; g.sz = P.u.sz; g.dx = gg.dx; g.dy = gg.dy; g.dz = gg.dz; g.dzc = gg.dzc; g.dzf = gg.dzf;
    // clamp out-of-range threads onto a valid column so that their (unused) pointers stay inside the arrays
    const int ii = in_range ? i : 1, jj = in_range ? j : 1;
    const long base = (long)ii + (long)jj * g.sy + (long)k0 * g.sz;  // every field has the same offsets on this path
    const T *pq = qf.p + qf.off + base;
    const T *pu = P.u.p + P.u.off + base, *pv = P.v.p + P.v.off + base, *pw = P.w.p + P.w.off + base;
    T lower = do_out ? fast_flux<T, N, FAST, WHICH, 2>(pq, pw, g, k0) : T(0);
    for (int k = k0; k <= k1; k++) {
        const T fx = do_x ? fast_flux<T, N, FAST, WHICH, 0>(pq, pu, g, k) : T(0);
        const T fy = do_y ? fast_flux<T, N, FAST, WHICH, 1>(pq, pv, g, k) : T(0);
        const T upper = do_out ? fast_flux<T, N, FAST, WHICH, 2>(pq + g.sz, pw + g.sz, g, k + 1) : T(0);
        const T fx1 = __shfl_down_sync(0xffffffffu, fx, 1);
        const int buf = k & 1;
        sy_buf[buf][ty][tx] = fy;
        __syncthreads();
        if (do_out) {
            const T fy1 = sy_buf[buf][ty + 1][tx];
            const T dzk = WHICH == 2 ? g.dzF(k) : g.dzC(k);
            const T Vi = 1 / ((g.dx * g.dy) * dzk);
            const T adv = Vi * ((fx1 - fx) + (fy1 - fy) + (upper - lower));
            T r;
            if constexpr (WHICH == 0) r = Gu_finish<T>(P, adv, i, j, k);
            else if constexpr (WHICH == 1) r = Gv_finish<T>(P, adv, i, j, k);
            else if constexpr (WHICH == 2) r = Gw_finish<T>(P, adv, i, j, k);
            else r = Gc_finish<T>(P, adv, t, i, j, k);
            G(i, j, k) = r;
        }
        lower = upper;
        pq += g.sz; pu += g.sz; pv += g.sz; pw += g.sz;
    }
}

// CTA-uniform predicate: every advective stencil of this k-chunk takes the full scheme
template <typename T, int N>
__device__ __forceinline__ bool fast_path_ok(const TendP<T> &P, int k0, int k1) {
    const GridD<T> &g = P.g;
    if (g.topo[0] != PERIODIC || g.topo[1] != PERIODIC) return false;
    if (g.topo[2] == PERIODIC) return true;
    if (g.topo[2] == FLAT) return false;
    // Bounded z: face-type flux indices k0 .. k1+1 and centre-type k0-1 .. k1 must satisfy outside_*_halo for buffer N
    return (k0 >= N + 1) && (k1 + 1 <= g.N[2] + 1 - N);
}

}  // namespace ob
