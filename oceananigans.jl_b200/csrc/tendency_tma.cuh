// tendency_tma.cuh -- TMA-staged variant of the interior fast path (tendency_fast.cuh).
//
// The marching kernel reads, per level k, the x- and y-lines of the advected field and of the advecting velocities from
// the planes k (and k+1 for w).  Every one of those loads is an LDG whose address is formed per thread; at 2N+2(N-1)
// values per flux and three fluxes per cell the LSU/address path and the L1 hit latency are what the FP64 pipe waits
// for (profiles/r1f notes).  Here ONE elected thread issues a `cp.async.bulk.tensor.3d` (TMA) per plane and level:
// the (32+2N) x (TY+2N) halo'd tile of u(k), v(k), w(k+1) (momentum) or of the advected field (w, tracers) lands in a
// three-slot shared-memory ring two levels ahead of its use, completion is signalled on an mbarrier, and the x/y lines
// become LDS with compile-time offsets.  z-lines (already perfectly coalesced, one value per thread and level) and the
// non-advective terms stay on the global path.  The arithmetic is flux_from_values() in both variants -- results are
// bit-identical to march_fast_body.
//
// Tensor maps: one per velocity component / tracer parent, dims (Px, Py, Pz), box (TW, TH, 1), no swizzle; built on
// the host for every launch (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, ~1 us each) and passed as a
// __grid_constant__ parameter.  Requires Px * sizeof(T) to be a multiple of 16 (checked on the host) and every box to
// start on a 16-byte boundary (the tile origin is rounded down in x; the box is widened by the same amount).
#pragma once
#include <cuda.h>
#include "tendency_fast.cuh"

namespace ob {

struct TmaMaps {
    CUtensorMap m[3 + OB_MAXTR];  // u, v, w, c[t]
};

template <typename T, int N, int TY>
struct TmaTile {
    static constexpr int EPV = 16 / (int)sizeof(T);                       // elements per 16 bytes
    static constexpr int TW = ((32 + 2 * N + 2 * (EPV - 1)) / EPV) * EPV;   // box width: 16-byte multiple, + the origin round-down
    static constexpr int TH = TY + 2 * N;
    static constexpr int BOX_BYTES = TW * TH * (int)sizeof(T);
    static constexpr int PLANE_BYTES = ((BOX_BYTES + 127) / 128) * 128;   // TMA destinations 128-byte aligned
    static constexpr int SLOTS = 3, NARR = 3;
    static constexpr int SMEM_BYTES = SLOTS * NARR * PLANE_BYTES;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// x-line / y-line of CNT values starting at offset lo from the thread's own point in a staged tile
template <int DIR, int CNT, int TW, typename T>
__device__ __forceinline__ void tile_line(const T *sb, int lo, T (&out)[CNT]) {
#pragma unroll
    for (int m = 0; m < CNT; m++) out[m] = sb[(lo + m) * (DIR == 0 ? 1 : TW)];
}

template <typename T, int N, bool FAST, int WHICH, int TY, int KC, bool STR>
__device__ __forceinline__ void march_tma_body(const TendP<T> &P, const TmaMaps &M, int t, int i0, int j0, int k0, int k1,
                                               T (*sy_buf)[TY][32], T (*sv_buf)[OB_SHARED_CL][TY][32], unsigned char *ring,
                                               uint64_t *bars) {
    using TT = TmaTile<T, N, TY>;
    constexpr int TW = TT::TW, NC = N - 1;
    constexpr int NARR = WHICH < 2 ? 3 : 1;   // momentum u, v: planes u(k), v(k), w(k+1); w and tracers: own plane k
    const GridD<T> &gg = P.g;
    const int Nx = gg.N[0], Ny = gg.N[1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = i0 + tx, j = j0 + ty;
    const bool do_out = (tx < 31) && (ty < TY - 1) && (i <= Nx) && (j <= Ny);
    const bool full_row = ty < TY - 1;
    const Fld<T> &qf = WHICH == 0 ? P.u : WHICH == 1 ? P.v : WHICH == 2 ? P.w : P.c[t];
    const Fld<T> &G = WHICH == 0 ? P.Gu : WHICH == 1 ? P.Gv : WHICH == 2 ? P.Gw : P.Gc[t];
    FastGeom<T, STR> g;
    g.init(gg, P.u.sy, P.u.sz);
    const int ii = min(i, Nx + 1), jj = min(j, Ny + 1);
    const int base = ii + jj * g.sy + k0 * g.sz;
    const T *pq = qf.p + qf.off + base;
    const T *pu = P.u.p + P.u.off + base, *pv = P.v.p + P.v.off + base, *pw = P.w.p + P.w.off + base;
    // staged tile: origin = logical (i0 - N, j0 - N); this thread's own point sits at (ii - i0 + N, jj - j0 + N)
    // (the box must start on a 16-byte boundary of the row: the x origin is rounded down to a multiple of EPV elements)
    const int cxu = i0 - N + gg.H[0] - 1;                                  // parent coordinates of the tile origin
    const int cx0 = cxu & ~(TT::EPV - 1), cy0 = j0 - N + gg.H[1] - 1, cz0 = k0 + gg.H[2] - 1;
    const int own = (jj - j0 + N) * TW + (ii - i0 + N) + (cxu - cx0);
    const bool producer = (tx == 0) && (ty == 0);
    const int nit = k1 - k0 + 1;
    const CUtensorMap *mq = &M.m[WHICH == 3 ? 3 + t : WHICH];
    auto issue = [&](int it) {
        const int s = it % TT::SLOTS;
        unsigned char *dst = ring + s * (TT::NARR * TT::PLANE_BYTES);
        mbar_expect_tx(&bars[s], NARR * TT::BOX_BYTES);
        if constexpr (WHICH < 2) {
            tma_load_3d(dst, &M.m[0], &bars[s], cx0, cy0, cz0 + it);
            tma_load_3d(dst + TT::PLANE_BYTES, &M.m[1], &bars[s], cx0, cy0, cz0 + it);
            tma_load_3d(dst + 2 * TT::PLANE_BYTES, &M.m[2], &bars[s], cx0, cy0, cz0 + it + 1);
        } else {
            tma_load_3d(dst, mq, &bars[s], cx0, cy0, cz0 + it);
        }
    };
    if (producer) {
        issue(0);
        if (nit > 1) issue(1);
    }
    const int ncl = P.ncl;
    const bool share_cl = ncl >= 1 && ncl <= OB_SHARED_CL;
    T lower = full_row ? fast_flux<T, N, FAST, WHICH, 2, STR>(pq, pw, g, k0) : T(0);
    T lower_c[OB_SHARED_CL] = {T(0), T(0)};
    if (share_cl && full_row) {
        FastTerms<T, STR> F0{P, g, pu, pv, pw, base, k0};
#pragma unroll
        for (int m = 0; m < OB_SHARED_CL; m++) if (m < ncl) lower_c[m] = F0.template first_lower_closure_flux<WHICH>(m, t, pq);
    }
    for (int k = k0; k <= k1; k++) {
        const int it = k - k0;
        const int s = it % TT::SLOTS;
        const T *tile = reinterpret_cast<const T *>(ring + s * (TT::NARR * TT::PLANE_BYTES)) + own;
        constexpr int PL = TT::PLANE_BYTES / (int)sizeof(T);
        const T *tq = tile + (WHICH == 1 ? PL : 0);   // the advected field's plane k
        mbar_wait(&bars[s], (uint32_t)((it / TT::SLOTS) & 1));
        T fx = T(0), upper = T(0), fy;
        {   // y flux: q along y; advecting v along axis WHICH
            T sq[2 * N];
            tile_line<1, 2 * N, TW>(tq, -N, sq);
            if constexpr (WHICH == 0) { T a[2 * NC]; tile_line<0, 2 * NC, TW>(tile + PL, -NC, a); fy = flux_from_values<T, N, FAST, WHICH, 1, STR>(sq, a, g, k); }
            else if constexpr (WHICH == 1) { T a[2 * NC]; tile_line<1, 2 * NC, TW>(tile + PL, -NC, a); fy = flux_from_values<T, N, FAST, WHICH, 1, STR>(sq, a, g, k); }
            else if constexpr (WHICH == 2) { T a[2 * NC]; load_line<2, 2 * NC>(pv, g, -NC, a); fy = flux_from_values<T, N, FAST, WHICH, 1, STR>(sq, a, g, k); }
            else { T a[1] = {__ldg(pv)}; fy = flux_from_values<T, N, FAST, WHICH, 1, STR>(sq, a, g, k); }
        }
        if (full_row) {
            {   // x flux
                T sq[2 * N];
                tile_line<0, 2 * N, TW>(tq, -N, sq);
                if constexpr (WHICH == 0) { T a[2 * NC]; tile_line<0, 2 * NC, TW>(tile, -NC, a); fx = flux_from_values<T, N, FAST, WHICH, 0, STR>(sq, a, g, k); }
                else if constexpr (WHICH == 1) { T a[2 * NC]; tile_line<1, 2 * NC, TW>(tile, -NC, a); fx = flux_from_values<T, N, FAST, WHICH, 0, STR>(sq, a, g, k); }
                else if constexpr (WHICH == 2) { T a[2 * NC]; load_line<2, 2 * NC>(pu, g, -NC, a); fx = flux_from_values<T, N, FAST, WHICH, 0, STR>(sq, a, g, k); }
                else { T a[1] = {__ldg(pu)}; fx = flux_from_values<T, N, FAST, WHICH, 0, STR>(sq, a, g, k); }
            }
            {   // upper z flux at k+1: q along z from global; advecting w(k+1) from the staged plane for u, v
                T sq[2 * N];
                load_line<2, 2 * N>(pq + g.sz, g, -N, sq);
                if constexpr (WHICH == 0) { T a[2 * NC]; tile_line<0, 2 * NC, TW>(tile + 2 * PL, -NC, a); upper = flux_from_values<T, N, FAST, WHICH, 2, STR>(sq, a, g, k + 1); }
                else if constexpr (WHICH == 1) { T a[2 * NC]; tile_line<1, 2 * NC, TW>(tile + 2 * PL, -NC, a); upper = flux_from_values<T, N, FAST, WHICH, 2, STR>(sq, a, g, k + 1); }
                else if constexpr (WHICH == 2) { T a[2 * NC]; load_line<2, 2 * NC>(pw + g.sz, g, -NC, a); upper = flux_from_values<T, N, FAST, WHICH, 2, STR>(sq, a, g, k + 1); }
                else { T a[1] = {__ldg(pw + g.sz)}; upper = flux_from_values<T, N, FAST, WHICH, 2, STR>(sq, a, g, k + 1); }
            }
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
        // every thread has finished level k-1 entirely and the staged reads of level k: slot (it+2)%3 == (it-1)%3 is free
        if (producer && it + 2 < nit) issue(it + 2);
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

}  // namespace ob
