// tendency_tiled.cuh -- the fused tendency kernel, flux-sharing marching form (the hot kernel of the time step).
//
// The reference evaluates every face flux twice (once per neighbouring cell, through δ(f) =  f(i+1) - f(i),
// src/Operators/difference_operators.jl:20-27) in 3+n separate launches.  Here ONE launch computes all 3+n
// tendencies and every advective face flux exactly ONCE:
//
//   * blockIdx.y selects the tendency (Gu, Gv, Gw, Gc[t]); each has exactly one x-, one y- and one z-direction
//     advective flux per cell, so every CTA runs the same shape of work and holds little state.
//   * a CTA is a 32 x TY tile of (i, j) columns that MARCHES in k over a chunk of KC levels:
//       - z-direction flux: kept in a register from one level to the next,
//       - x-direction flux: taken from lane+1 with a warp shuffle (lane 31 is the overlap lane: tiles advance by 31),
//       - y-direction flux: taken from row+1 through double-buffered shared memory (one __syncthreads per level;
//         row TY-1 is the overlap row and computes only its y-flux: tiles advance by TY-1).
//   * the flux functions are the SAME device functions the generic kernel uses (mom_flux / tracer_flux with the
//     Bounded fallback chain and Flat rules), so the arithmetic of every flux is unchanged; the non-advective terms
//     (buoyancy, Coriolis, hydrostatic pressure gradient, closures) are added by the shared G*_finish functions.
//
// Inputs are read with ld.global.nc through L1 (coalesced along x); the kernel is FP64-issue bound, not HBM or
// L1 bound, so no shared-memory staging of the input tile is needed.
#pragma once
#include <vector>
#include <stdlib.h>
#include "tendency.cuh"
#include "tendency_tma.cuh"
#include "tma_maps.h"
#include "stage_launch.h"

namespace ob {

// own-direction advective flux of tendency WHICH (0 u, 1 v, 2 w, 3 tracer) at flux index (i, j, k)
template <typename T, class S, bool FAST, int WHICH, int DIR>
__device__ __forceinline__ T adv_flux(const TendP<T> &P, const Fld<T> &c, int i, int j, int k) {
    if constexpr (WHICH == 3) return tracer_flux<T, S, FAST, DIR>(P, c, i, j, k);
    else return mom_flux<T, S, FAST, DIR, WHICH>(P, i, j, k);
}

template <typename T, class S, bool FAST, int WHICH, int TY, int KC>
__device__ __forceinline__ void march_body(const TendP<T> &P, int t, int i, int j, int k0, int k1, T (*sy)[TY][32]) {
    const GridD<T> &g = P.g;
    const int Nx = g.N[0], Ny = g.N[1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    // centre-type directions: the tendency at index n uses F(n) - F(n-1), so the thread at n owns F(n-1)
    constexpr int cx = WHICH == 0, cy = WHICH == 1, cz = WHICH == 2;
    const bool do_x = (ty < TY - 1) && (j <= Ny) && (i <= Nx + 1);
    const bool do_y = (tx < 31) && (i <= Nx) && (j <= Ny + 1);
    const bool do_out = (tx < 31) && (ty < TY - 1) && (i <= Nx) && (j <= Ny);
    const Fld<T> &c = P.c[WHICH == 3 ? t : 0];
    const Fld<T> &G = WHICH == 0 ? P.Gu : WHICH == 1 ? P.Gv : WHICH == 2 ? P.Gw : P.Gc[t];
    // volume at the tendency location: V = (Δx Δy) Δz (spacings_and_areas_and_volumes.jl:483-491)
    T lower = do_out ? adv_flux<T, S, FAST, WHICH, 2>(P, c, i, j, k0 - cz) : T(0);
    for (int k = k0; k <= k1; k++) {
        const T fx = do_x ? adv_flux<T, S, FAST, WHICH, 0>(P, c, i - cx, j, k) : T(0);
        const T fy = do_y ? adv_flux<T, S, FAST, WHICH, 1>(P, c, i, j - cy, k) : T(0);
        const T upper = do_out ? adv_flux<T, S, FAST, WHICH, 2>(P, c, i, j, k + 1 - cz) : T(0);
        const T fx1 = __shfl_down_sync(0xffffffffu, fx, 1);
        const int buf = k & 1;
        sy[buf][ty][tx] = fy;
        __syncthreads();
        if (do_out) {
            const T fy1 = sy[buf][ty + 1][tx];
            const T Vi = WHICH == 2 ? g.rVf(k) : g.rVc(k);
            const T ddx = g.topo[0] == FLAT ? T(0) : fx1 - fx;
            const T ddy = g.topo[1] == FLAT ? T(0) : fy1 - fy;
            const T ddz = g.topo[2] == FLAT ? T(0) : upper - lower;
            const T adv = Vi * (ddx + ddy + ddz);
            T r;
            if constexpr (WHICH == 0) r = Gu_finish<T>(P, adv, i, j, k);
            else if constexpr (WHICH == 1) r = Gv_finish<T>(P, adv, i, j, k);
            else if constexpr (WHICH == 2) r = Gw_finish<T>(P, adv, i, j, k);
            else r = Gc_finish<T>(P, adv, t, i, j, k);
            G(i, j, k) = r;
        }
        lower = upper;
    }
}

template <typename T, class S, bool FAST, int TY, int KC, int MINB>
__global__ void __launch_bounds__(32 * TY, MINB) tendency_march_kernel(const __grid_constant__ TendP<T> P, int nb, int nkc, int tx_lo, int ntx, int skip_lo, int skip_n, int wall_only) {
    __shared__ T sy[2][TY][32];
    __shared__ T sv[2][OB_SHARED_CL][TY][32];
    const int which = blockIdx.y;
    const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
    const int nty = (Ny + TY - 2) / (TY - 1);   // x tiles tx_lo .. tx_lo + ntx - 1 of this launch (all of them unless split)
    const int b = blockIdx.x;
    int tile_x = tx_lo + b % ntx;
    if (tile_x >= skip_lo) tile_x += skip_n;   // complement launches: every tile except [skip_lo, skip_lo + skip_n)
    const int tile_y = (b / ntx) % nty;
    int kc = b / (ntx * nty);
    if (wall_only) kc = kc ? nkc - 1 : 0;      // only the two wall-adjacent chunks (the staged-ring kernel does the interior)
    const int i = 1 + tile_x * 31 + (int)threadIdx.x, j = 1 + tile_y * (TY - 1) + (int)threadIdx.y;
    // k-chunks: with a Bounded z and a WENO scheme of buffer nb the first and last chunk are the nb wall-adjacent
    // levels (general path with the fallback chain); every other chunk is interior and takes the fast path
    int k0, k1;
    if (nb == 0) { k0 = 1 + kc * KC; k1 = min(k0 + KC - 1, Nz); }
    else if (kc == 0) { k0 = 1; k1 = nb; }
    else if (kc == nkc - 1) { k0 = Nz - nb + 1; k1 = Nz; }
    else { k0 = nb + 1 + (kc - 1) * KC; k1 = min(k0 + KC - 1, Nz - nb); }
    if constexpr (S::kind == ADV_WENO) {
        if (fast_path_ok<T, S::n>(P, k0, k1)) {  // CTA-uniform
            const int t = which - 3;
            if (P.g.dzc) {
                if (which == 0) march_fast_body<T, S::n, FAST, 0, TY, KC, true>(P, 0, i, j, k0, k1, sy, sv);
                else if (which == 1) march_fast_body<T, S::n, FAST, 1, TY, KC, true>(P, 0, i, j, k0, k1, sy, sv);
                else if (which == 2) march_fast_body<T, S::n, FAST, 2, TY, KC, true>(P, 0, i, j, k0, k1, sy, sv);
                else march_fast_body<T, S::n, FAST, 3, TY, KC, true>(P, t, i, j, k0, k1, sy, sv);
            } else {
                if (which == 0) march_fast_body<T, S::n, FAST, 0, TY, KC, false>(P, 0, i, j, k0, k1, sy, sv);
                else if (which == 1) march_fast_body<T, S::n, FAST, 1, TY, KC, false>(P, 0, i, j, k0, k1, sy, sv);
                else if (which == 2) march_fast_body<T, S::n, FAST, 2, TY, KC, false>(P, 0, i, j, k0, k1, sy, sv);
                else march_fast_body<T, S::n, FAST, 3, TY, KC, false>(P, t, i, j, k0, k1, sy, sv);
            }
            return;
        }
    }
    if (which == 0) march_body<T, S, FAST, 0, TY, KC>(P, 0, i, j, k0, k1, sy);
    else if (which == 1) march_body<T, S, FAST, 1, TY, KC>(P, 0, i, j, k0, k1, sy);
    else if (which == 2) march_body<T, S, FAST, 2, TY, KC>(P, 0, i, j, k0, k1, sy);
    else march_body<T, S, FAST, 3, TY, KC>(P, which - 3, i, j, k0, k1, sy);
}

template <typename T, class S, int TY, int KC, int MINB>
static cudaError_t launch_march(const TendP<T> &P, int fast, cudaStream_t st, int *nlaunch, int tx_lo, int tx_hi, int invert, int wall_only = 0) {
    const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
    int nb = 0;
    long nkc = (Nz + KC - 1) / KC;
    if (S::kind == ADV_WENO && P.g.topo[2] == BOUNDED && P.g.topo[0] == PERIODIC && P.g.topo[1] == PERIODIC && Nz > 2 * S::n) {
        nb = S::n;
        nkc = 2 + (Nz - 2 * nb + KC - 1) / KC;
    }
    const long ntx_all = (Nx + OB_TILE_X - 1) / OB_TILE_X, nty = (Ny + TY - 2) / (TY - 1);
    if (tx_hi < 0 || tx_hi > ntx_all) tx_hi = (int)ntx_all;
    int skip_lo = 1 << 30, skip_n = 0;
    long ntx = tx_hi - tx_lo;
    if (invert) { skip_lo = tx_lo; skip_n = tx_hi - tx_lo; ntx = ntx_all - skip_n; tx_lo = 0; }
    if (ntx <= 0) return cudaSuccess;
    if (ntx * nty * nkc > 2147483647L) return cudaErrorInvalidConfiguration;
    if (wall_only && nb == 0) return cudaErrorInvalidValue;
    dim3 grid((unsigned)(ntx * nty * (wall_only ? 2 : nkc)), 3 + P.ntr), block(32, TY);
    if (S::kind == ADV_WENO && fast) tendency_march_kernel<T, S, true, TY, KC, MINB><<<grid, block, 0, st>>>(P, nb, (int)nkc, tx_lo, (int)ntx, skip_lo, skip_n, wall_only);
    else tendency_march_kernel<T, S, false, TY, KC, MINB><<<grid, block, 0, st>>>(P, nb, (int)nkc, tx_lo, (int)ntx, skip_lo, skip_n, wall_only);
    *nlaunch += 1;
    return cudaGetLastError();
}

// ---- TMA-staged variant (tendency_tma.cuh) ------------------------------------------------------------------------
template <typename T, class S, bool FAST, int TY, int KC, int MINB>
__global__ void __launch_bounds__(32 * TY, MINB) tendency_march_tma_kernel(const __grid_constant__ TendP<T> P, const __grid_constant__ TmaMaps M,
                                                                          int nb, int nkc, int tx_lo, int ntx, int skip_lo, int skip_n) {
    __shared__ T sy[2][TY][32];
    __shared__ T sv[2][OB_SHARED_CL][TY][32];
    __shared__ __align__(8) uint64_t bars[3];
    extern __shared__ __align__(128) unsigned char ring[];
    const int which = blockIdx.y;
    const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
    const int nty = (Ny + TY - 2) / (TY - 1);   // x tiles tx_lo .. tx_lo + ntx - 1 of this launch (all of them unless split)
    const int b = blockIdx.x;
    int tile_x = tx_lo + b % ntx;
    if (tile_x >= skip_lo) tile_x += skip_n;   // complement launches: every tile except [skip_lo, skip_lo + skip_n)
    const int tile_y = (b / ntx) % nty, kc = b / (ntx * nty);
    const int i0 = 1 + tile_x * 31, j0 = 1 + tile_y * (TY - 1);
    const int i = i0 + (int)threadIdx.x, j = j0 + (int)threadIdx.y;
    int k0, k1;
    if (nb == 0) { k0 = 1 + kc * KC; k1 = min(k0 + KC - 1, Nz); }
    else if (kc == 0) { k0 = 1; k1 = nb; }
    else if (kc == nkc - 1) { k0 = Nz - nb + 1; k1 = Nz; }
    else { k0 = nb + 1 + (kc - 1) * KC; k1 = min(k0 + KC - 1, Nz - nb); }
    if constexpr (S::kind == ADV_WENO) {
        if (fast_path_ok<T, S::n>(P, k0, k1)) {  // CTA-uniform
            if (threadIdx.x == 0 && threadIdx.y == 0) {
                for (int s = 0; s < 3; s++) mbar_init(&bars[s], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            const int t = which - 3;
            if (P.g.dzc) {
                if (which == 0) march_tma_body<T, S::n, FAST, 0, TY, KC, true>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else if (which == 1) march_tma_body<T, S::n, FAST, 1, TY, KC, true>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else if (which == 2) march_tma_body<T, S::n, FAST, 2, TY, KC, true>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else march_tma_body<T, S::n, FAST, 3, TY, KC, true>(P, M, t, i0, j0, k0, k1, sy, sv, ring, bars);
            } else {
                if (which == 0) march_tma_body<T, S::n, FAST, 0, TY, KC, false>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else if (which == 1) march_tma_body<T, S::n, FAST, 1, TY, KC, false>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else if (which == 2) march_tma_body<T, S::n, FAST, 2, TY, KC, false>(P, M, 0, i0, j0, k0, k1, sy, sv, ring, bars);
                else march_tma_body<T, S::n, FAST, 3, TY, KC, false>(P, M, t, i0, j0, k0, k1, sy, sv, ring, bars);
            }
            return;
        }
    }
    if (which == 0) march_body<T, S, FAST, 0, TY, KC>(P, 0, i, j, k0, k1, sy);
    else if (which == 1) march_body<T, S, FAST, 1, TY, KC>(P, 0, i, j, k0, k1, sy);
    else if (which == 2) march_body<T, S, FAST, 2, TY, KC>(P, 0, i, j, k0, k1, sy);
    else march_body<T, S, FAST, 3, TY, KC>(P, which - 3, i, j, k0, k1, sy);
}

// true when the TMA variant applies: WENO, (Periodic, Periodic, non-Flat), 16-byte row pitch, one parent shape per field
template <typename T, class S>
static bool tma_applicable(const TendP<T> &P) {
    if (S::kind != ADV_WENO) return false;
    if (P.g.topo[0] != PERIODIC || P.g.topo[1] != PERIODIC || P.g.topo[2] == FLAT) return false;
    if ((P.u.sy * sizeof(T)) % 16 != 0) return false;
    return encode_tiled_fn() != nullptr;
}

template <typename T, class S, int TY, int KC, int MINB>
static cudaError_t launch_march_tma(const TendP<T> &P, int fast, cudaStream_t st, int *nlaunch, int tx_lo, int tx_hi, int invert) {
    if constexpr (S::kind != ADV_WENO) return cudaErrorNotSupported;
    else {
        using TT = TmaTile<T, S::n, TY>;
        const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
        int nb = 0;
        long nkc = (Nz + KC - 1) / KC;
        if (P.g.topo[2] == BOUNDED && Nz > 2 * S::n) { nb = S::n; nkc = 2 + (Nz - 2 * nb + KC - 1) / KC; }
        const long ntx_all = (Nx + OB_TILE_X - 1) / OB_TILE_X, nty = (Ny + TY - 2) / (TY - 1);
        if (tx_hi < 0 || tx_hi > ntx_all) tx_hi = (int)ntx_all;
        int skip_lo = 1 << 30, skip_n = 0;
        long ntx = tx_hi - tx_lo;
        if (invert) { skip_lo = tx_lo; skip_n = tx_hi - tx_lo; ntx = ntx_all - skip_n; tx_lo = 0; }
        if (ntx <= 0) return cudaSuccess;
        if (ntx * nty * nkc > 2147483647L) return cudaErrorInvalidConfiguration;
        TmaMaps M;
        memset(&M, 0, sizeof(M));
        const cuuint64_t Px = (cuuint64_t)P.u.sy, Py = (cuuint64_t)(P.u.sz / P.u.sy);
        auto encode = [&](CUtensorMap *m, const Fld<T> &f, bool face_z) -> bool {
            const cuuint64_t Pz = (cuuint64_t)(Nz + 2 * P.g.H[2] + ((face_z && P.g.topo[2] == BOUNDED) ? 1 : 0));
            const cuuint64_t dims[3] = {Px, Py, Pz};
            const cuuint64_t strides[2] = {Px * sizeof(T), Px * Py * sizeof(T)};
            const cuuint32_t box[3] = {(cuuint32_t)TT::TW, (cuuint32_t)TT::TH, 1u};
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            return encode_tiled_fn()(m, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)f.p, dims,
                                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        bool ok = encode(&M.m[0], P.u, false) && encode(&M.m[1], P.v, false) && encode(&M.m[2], P.w, true);
        for (int t = 0; ok && t < P.ntr; t++) ok = encode(&M.m[3 + t], P.c[t], false);
        if (!ok) return cudaErrorInvalidValue;
        dim3 grid((unsigned)(ntx * nty * nkc), 3 + P.ntr), block(32, TY);
        cudaError_t e;
        if (fast) {
            auto kern = tendency_march_tma_kernel<T, S, true, TY, KC, MINB>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TT::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            kern<<<grid, block, TT::SMEM_BYTES, st>>>(P, M, nb, (int)nkc, tx_lo, (int)ntx, skip_lo, skip_n);
        } else {
            auto kern = tendency_march_tma_kernel<T, S, false, TY, KC, MINB>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TT::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            kern<<<grid, block, TT::SMEM_BYTES, st>>>(P, M, nb, (int)nkc, tx_lo, (int)ntx, skip_lo, skip_n);
        }
        *nlaunch += 1;
        return cudaGetLastError();
    }
}

// which (float type, scheme) combinations have a staged-ring kernel (stage_launch.h lists the variants)
template <typename T, class S>
struct StageSel {
    static constexpr bool built = S::kind == ADV_WENO && S::n == 3;
};

// ---- staged-ring kernel (tendency_stage.cuh, instantiated in stage_inst.cu) ------------------------------------------------
// Applicable: WENO, (Periodic, Periodic, Periodic | Bounded) topology, the rcp + Newton division, at most two closures, none
// vertically implicit, 32-bit element offsets, the whole x range in one launch; closures either all ScalarDiffusivity (mode MT)
// or ONE eddy-viscosity closure in first position followed by at most one ScalarDiffusivity (modes MN + TT).
template <typename T, class S>
static bool stage_applicable(const TendP<T> &P, int &les_kind) {
    les_kind = 0;
    if (S::kind != ADV_WENO) return false;
    const GridD<T> &g = P.g;
    if (g.topo[0] != PERIODIC || g.topo[1] != PERIODIC || g.topo[2] == FLAT) return false;
    if (P.ncl > OB_SHARED_CL) return false;
    for (int m = 0; m < P.ncl; m++) if (P.cl[m].vi) return false;
    for (int m = 1; m < P.ncl; m++) if (P.cl[m].kind != CL_SCALAR) return false;
    if (P.ncl >= 1 && P.cl[0].kind != CL_SCALAR) les_kind = P.cl[0].kind;
    if (P.ncl == 2 && les_kind == 0) return false;   // two ScalarDiffusivities: not instantiated (marching kernel)
    if (P.u.sz * (long)(g.N[2] + 2 * g.H[2] + 1) >= 2147483647L) return false;
    if (g.topo[2] == BOUNDED && g.N[2] <= 2 * S::n) return false;
    for (int d = 0; d < 3; d++) if (g.H[d] < S::n) return false;
    return true;
}

// mode: 0 auto, 1 generic, 2 marching (LDG), 3 marching with TMA-staged planes (4.. = tuning variants when built with -DOB_TI_EXPERIMENT)
template <typename T, class S>
static cudaError_t try_tiled_tendency(const TendP<T> &P, int fast, int mode, cudaStream_t st, int sm_count, int *nlaunch, bool &done, int tx_lo, int tx_hi, int invert) {
    (void)sm_count;
    done = false;
    if (mode == 1) return cudaSuccess;
    // WENO-11: the one-thread-per-cell kernel only (the marching form of a 12-point stencil spills 10 KB per thread and takes
    // seven minutes to compile; the scheme is outside BASELINE.json's configurations)
    if constexpr (S::kind == ADV_WENO && S::n >= 6) return cudaSuccess;
    else {
    done = true;
    const bool whole = tx_lo == 0 && tx_hi < 0 && !invert;
    if constexpr (StageSel<T, S>::built) {
        int les_kind = 0;
        if ((mode == 0 || mode == 8) && whole && fast && stage_applicable<T, S>(P, les_kind)) {
            cudaError_t e;
            if (les_kind == 0) {
                e = launch_stage_variant(P, S::n, STAGE_MODE_MT, P.ncl, 0, st, sm_count, nlaunch);
            } else {
                e = launch_stage_variant(P, S::n, STAGE_MODE_MN, P.ncl, les_kind, st, sm_count, nlaunch);
                if (e == cudaSuccess && P.ntr > 0) e = launch_stage_variant(P, S::n, STAGE_MODE_TT, P.ncl, les_kind, st, sm_count, nlaunch);
            }
            if (e != cudaSuccess) return e;
            // Bounded z: the wall-adjacent chunks take the marching kernel with the fallback chain
            if (P.g.topo[2] == BOUNDED) return launch_march<T, S, 8, 32, 4>(P, 1, st, nlaunch, 0, -1, 0, 1);
            return cudaSuccess;
        }
    }
    if (mode == 8) mode = 0;
    // auto: Float32 takes the TMA-staged variant (measured 4.5 % faster at 256^3), Float64 the LDG one (2 % faster)
    if (mode == 0 && sizeof(T) == 4 && S::kind == ADV_WENO) mode = 3;
    if (mode == 3) {   // TMA-staged planes (falls back to the LDG marching kernel where TMA does not apply)
        if (tma_applicable<T, S>(P)) return launch_march_tma<T, S, 8, 32, 4>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
        return launch_march<T, S, 8, 32, 4>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
    }
#ifdef OB_TI_EXPERIMENT
    if (mode == 4) return launch_march<T, S, 4, 32, 8>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
    if (mode == 5) return launch_march<T, S, 16, 32, 2>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
    if (mode == 6) return launch_march<T, S, 8, 32, 3>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
    if (mode == 7) return launch_march<T, S, 8, 16, 4>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
#endif
    return launch_march<T, S, 8, 32, 4>(P, fast, st, nlaunch, tx_lo, tx_hi, invert);
    }
}

}  // namespace ob
