// streaming.cuh -- 128-bit forms of the bandwidth-bound kernels of a substep (update + tendency cache, Poisson source term,
// fused projection).  Same arithmetic, expression for expression, as the one-cell-per-thread kernels of kernels.cuh
// (tests/test_gpu_parity.py compares the two bit for bit); what changes is the access pattern:
//
//   * a thread owns a 16-byte-aligned group of V consecutive x elements of the PARENT array (V = 2 doubles / 4 floats in the
//     update kernel, one pair in the stencil kernels) and ROWS consecutive rows, so every field access of the update is one
//     128-bit load / store and 2 x ROWS of them are in flight per thread before the first use;
//   * the interior of a row starts Hx elements into the parent row, i.e. generally NOT on a 16-byte boundary: the group that
//     straddles the west (east) end of the interior is loaded whole -- its extra lanes are halo cells of the same row, or the
//     tail of the neighbouring row, always inside the allocation -- and stored lane by lane.  The alignment phase is computed
//     per row from the element offset, so no assumption is made on the row pitch (Float32 at Nx = 256, H = 3 has a
//     1048-byte pitch: rows alternate between two phases).
//
// The host checks what the kernels assume (vector_ok() in ocean_b200.cu): 16-byte-aligned base pointers, the same alignment
// phase for every field one thread touches, at least one halo row in front of the first interior row.
// Reference kernels restated: runge_kutta_3.jl:196-204, quasi_adams_bashforth_2.jl:134-147,
// cache_nonhydrostatic_tendencies.jl:8-31, solve_for_pressure.jl:12-42, pressure_correction.jl:67-103.
#pragma once
#include "kernels.cuh"

namespace ob {

template <typename T> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int V = 2; };
template <> struct Vec16<float> { using type = float4; static constexpr int V = 4; };
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename VT, typename T, int V>
__device__ __forceinline__ void vec_load(T (&dst)[V], const T *p) { *reinterpret_cast<VT *>(dst) = *reinterpret_cast<const VT *>(p); }
template <typename VT, typename T, int V>
__device__ __forceinline__ void vec_store(T *p, const T (&src)[V]) { *reinterpret_cast<VT *>(p) = *reinterpret_cast<const VT *>(src); }

// blockIdx.x enumerates (x block of 128 groups, block of ROWS rows, level); blockIdx.y the field
template <int ROWS>
__device__ __forceinline__ bool group_from_block(int ngx, int Ny, int &t, int &j0, int &k) {
    const int nbx = (ngx + 127) / 128, njb = (Ny + ROWS - 1) / ROWS;
    long b = blockIdx.x;
    const int bx = (int)(b % nbx); b /= nbx;
    j0 = 1 + (int)(b % njb) * ROWS;
    k = 1 + (int)(b / njb);
    t = bx * 128 + (int)threadIdx.x;
    return t < ngx;
}

// ------------------------------------------------------------------------------------------------------------
// update_kernel, vector form.  ngx = groups per row = (Nx + V - 1) / V + 1.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int ROWS>
__global__ void __launch_bounds__(128) update_vec_kernel(const __grid_constant__ UpdateP<T> P, int ngx) {
    using VT = typename Vec16<T>::type;
    constexpr int V = Vec16<T>::V;
    const int f = blockIdx.y;
    int t, j0, k;
    if (!group_from_block<ROWS>(ngx, P.N[1], t, j0, k)) return;
    const Fld<T> U = P.U[f], Gn = P.Gn[f], Gm = P.Gm[f];
    const int Nx = P.N[0], Ny = P.N[1];
    const int lo0 = P.lo[f][0];
    const bool klive = k >= P.lo[f][2];
    const bool need_gm = P.mode == 1 || (P.mode == 2 && P.chi != T(-0.5));
    T u[ROWS][V], gn[ROWS][V], gm[ROWS][V];
    long eu[ROWS], eg[ROWS], em[ROWS];
    int i0[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = min(j0 + r, Ny);   // (rows past Ny re-read the last row and store nothing)
        const long bu = U.off + (long)j * U.sy + (long)k * U.sz;
        i0[r] = V * t - (int)(bu & (V - 1));   // logical i of lane 0: (bu + i0) is a multiple of V
        eu[r] = bu + i0[r];
        eg[r] = Gn.off + (long)j * Gn.sy + (long)k * Gn.sz + i0[r];
        em[r] = Gm.off + (long)j * Gm.sy + (long)k * Gm.sz + i0[r];
        vec_load<VT>(u[r], U.p + eu[r]);
        vec_load<VT>(gn[r], Gn.p + eg[r]);
        if (need_gm) vec_load<VT>(gm[r], Gm.p + em[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = j0 + r;
        if (j > Ny) break;
        const bool jlive = klive && j >= P.lo[f][1];
        bool all_valid = true, all_active = true, any_active = false;
        bool act[V];
#pragma unroll
        for (int l = 0; l < V; l++) {
            const int i = i0[r] + l;
            const bool valid = i >= 1 && i <= Nx;
            act[l] = valid && jlive && i >= lo0;
            all_valid &= valid; all_active &= act[l]; any_active |= act[l];
            u[r][l] = updated_value(P, u[r][l], gn[r][l], need_gm ? gm[r][l] : T(0));
        }
        if (all_active) vec_store<VT>(U.p + eu[r], u[r]);
        else if (any_active) {
#pragma unroll
            for (int l = 0; l < V; l++) if (act[l]) U.p[eu[r] + l] = u[r][l];
        }
        if (P.do_cache) {
            if (all_valid) vec_store<VT>(Gm.p + em[r], gn[r]);
            else {
#pragma unroll
                for (int l = 0; l < V; l++) { const int i = i0[r] + l; if (i >= 1 && i <= Nx) Gm.p[em[r] + l] = gn[r][l]; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// source_term_kernel, pair form: a thread owns the aligned pair (i0, i0+1) of u / v / w in ROWS rows.  The row pitch and the
// level pitch of the velocity parents are even (checked by the host), so the pairs of rows j+1 and levels k+1 are aligned too.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int ROWS>
__global__ void __launch_bounds__(128) source_pair_kernel(const __grid_constant__ SourceP<T> P, int ngx) {
    using VT = typename Vec2<T>::type;
    int t, j0, k;
    const int Nx = P.g.N[0], Ny = P.g.N[1];
    if (!group_from_block<ROWS>(ngx, Ny, t, j0, k)) return;
    const T dzc = P.g.dzC(k);
    const T Ax = P.g.dy * dzc, Ay = P.g.dx * dzc, Az = P.g.dx * P.g.dy;
    const T Vi = P.g.rVc(k);
    const int kk = P.zperm ? makhoul_index(k - 1, P.g.N[2]) : k - 1;
    T u[ROWS][2], ue[ROWS], v[ROWS + 1][2], w0[ROWS][2], w1[ROWS][2];
    // the phase is the same for every row and for u, v, w (even pitches, same parity of `off`: host-checked)
    const int i0 = 2 * t - (int)((P.u.off + (long)j0 * P.u.sy + (long)k * P.u.sz) & 1);
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = min(j0 + r, Ny);
        vec_load<VT>(u[r], P.u.p + P.u.idx(i0, j, k));
        ue[r] = P.u.ld(min(i0 + 2, Nx + 1), j, k);
        vec_load<VT>(w0[r], P.w.p + P.w.idx(i0, j, k));
        vec_load<VT>(w1[r], P.w.p + P.w.idx(i0, j, k + 1));
    }
#pragma unroll
    for (int r = 0; r <= ROWS; r++) vec_load<VT>(v[r], P.v.p + P.v.idx(i0, min(j0 + r, Ny + 1), k));
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = j0 + r;
        if (j > Ny) break;
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int i = i0 + l;
            if (i < 1 || i > Nx) continue;
            const T div = source_value(P, Ax, Ay, Az, Vi, dzc, u[r][l], l == 0 ? u[r][1] : ue[r], v[r][l], v[r + 1][l], w0[r][l], w1[r][l]);
            const long o = (i - 1) + (j - 1) * P.ldx + (long)kk * P.ldxy;
            if (P.cplx) { P.out[2 * o] = div; P.out[2 * o + 1] = T(0); }
            else P.out[o] = div;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// correct_fused_kernel, pair form: u, v, w are read and written as aligned pairs, p stored as a pair; the solver output is
// read through the read-only path (its x neighbours are mostly the thread's own or its neighbour's values).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int ROWS>
__global__ void __launch_bounds__(128) correct_pair_kernel(const __grid_constant__ CorrectFusedP<T> P, int ngx) {
    using VT = typename Vec2<T>::type;
    int t, j0, k;
    const int Nx = P.g.N[0], Ny = P.g.N[1];
    if (!group_from_block<ROWS>(ngx, Ny, t, j0, k)) return;
    auto S = [&](int a, int b, int c) -> T {
        const int cc = P.zperm ? makhoul_index(c - 1, P.g.N[2]) : c - 1;
        const long o = (a - 1) + (b - 1) * P.ldx + (long)cc * P.ldxy;
        return mul_rn(P.cplx ? __ldg(P.sol + 2 * o) : __ldg(P.sol + o), P.scale);
    };
    auto lower = [&](int idx, int d) { return (idx > 1 || (d == 0 && P.west)) ? idx - 1 : (P.g.topo[d] == PERIODIC ? P.g.N[d] : 1); };
    const int i0 = 2 * t - (int)((P.u.off + (long)j0 * P.u.sy + (long)k * P.u.sz) & 1);
    const T rdzf = P.g.rdzF(k);
    const int kl = lower(k, 2);
    T u[ROWS][2], v[ROWS][2], w[ROWS][2];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = min(j0 + r, Ny);
        vec_load<VT>(u[r], P.u.p + P.u.idx(i0, j, k));
        vec_load<VT>(v[r], P.v.p + P.v.idx(i0, j, k));
        vec_load<VT>(w[r], P.w.p + P.w.idx(i0, j, k));
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int j = j0 + r;
        if (j > Ny) break;
        const int jl = lower(j, 1);
        T pc[2], pw[2], ps[2], pb[2];
        bool ok[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
            const int i = i0 + l;
            ok[l] = i >= 1 && i <= Nx;
            const int ic = ok[l] ? i : 1;
            pc[l] = S(ic, j, k);
            ps[l] = S(ic, jl, k);
            pb[l] = S(ic, j, kl);
        }
        pw[0] = S(lower(ok[0] ? i0 : 1, 0), j, k);
        pw[1] = ok[0] ? pc[0] : S(lower(1, 0), j, k);   // (lane 1 at i = 1: lane 0 is the halo cell)
        T p[2];
#pragma unroll
        for (int l = 0; l < 2; l++) {
            u[r][l] = sub_rn(u[r][l], mul_rn(sub_rn(pc[l], pw[l]), P.g.rdx));
            v[r][l] = sub_rn(v[r][l], mul_rn(sub_rn(pc[l], ps[l]), P.g.rdy));
            w[r][l] = sub_rn(w[r][l], mul_rn(sub_rn(pc[l], pb[l]), rdzf));
            p[l] = pc[l] / P.denom;
        }
        if (ok[0] && ok[1]) {
            vec_store<VT>(P.u.p + P.u.idx(i0, j, k), u[r]);
            vec_store<VT>(P.v.p + P.v.idx(i0, j, k), v[r]);
            vec_store<VT>(P.w.p + P.w.idx(i0, j, k), w[r]);
            vec_store<VT>(P.p.p + P.p.idx(i0, j, k), p);
        } else {
#pragma unroll
            for (int l = 0; l < 2; l++)
                if (ok[l]) { P.u(i0 + l, j, k) = u[r][l]; P.v(i0 + l, j, k) = v[r][l]; P.w(i0 + l, j, k) = w[r][l]; P.p(i0 + l, j, k) = p[l]; }
        }
    }
}

}  // namespace ob
