// tendency_stage.cuh -- the fused tendency kernel, staged-ring form (default on the interior fast path).
//
// One CTA computes ALL tendencies of a 32 x W tile of columns (Gu, Gv, Gw and one tracer per pass) while it marches
// in k.  Every stencil operand is read from shared memory:
//
//   * a ring of D = N+2 levels holds the halo'd (32+2N) x (W+2N) planes of u, v, w and the tracer, brought in by TMA
//     (cp.async.bulk.tensor.3d, one box per field and level, completion on an mbarrier); the warp that is LAST to release
//     a level issues the loads of the level that takes its place -- no thread spins as a producer.  When the row pitch of
//     the parent arrays is not a multiple of 16 bytes (Float32 with odd padding) cp.async copies take the place of TMA.
//     Levels k .. k+N live in the ring; the N-1 levels below k of a thread's own column are kept in registers, so the
//     z-lines of the stencils cost no memory traffic at all.
//   * W COMPUTE warps: warp w owns row w of the tile, lane l the column i0+l.  A thread evaluates, per level and
//     tendency, the flux through the west face and the south face of its cell and through the upper face; the east flux
//     comes from lane+1 (shuffle), the north flux is the south flux the next warp published in shared memory, the lower
//     flux is last level's upper flux (register).  Every face flux is evaluated ONCE.
//   * four HELPER warps (one per SM sub-partition, so the FP64 pipes stay balanced) evaluate the tile's east-edge x
//     fluxes (lanes = rows) and north-edge y fluxes (lanes = columns) -- helper q those of tendency q -- and publish them,
//     so tiles advance by the full 32 x W cells: no overlap lanes, no overlap rows.  When TMA is not applicable the
//     helpers also feed the ring (helper f copies field f with cp.async).
//   * warps never meet at a CTA barrier: ring slots are recycled through full/empty mbarriers (the empty barrier of a
//     level needs one arrival per consumer warp), published fluxes through one mbarrier per publishing warp.  Because a
//     level k+N can only be loaded after EVERY consumer has released level k-2, two consumer warps are never more than one
//     level apart, which is what makes two publication slots sufficient.
//
// The arithmetic is flux_from_values() / the closure-flux expressions of tendency_fast.cuh, operand for operand, so the
// results are bit-identical to march_fast_body (tests/test_gpu_parity.py checks this and the oracle comparison).
//
// Reference semantics: compute_nonhydrostatic_tendencies.jl:107-150 (one kernel per tendency there),
// nonhydrostatic_tendency_kernel_functions.jl:71-302, upwind_biased_advective_fluxes.jl:23-121.
#pragma once
#include <cuda.h>
#include "tendency_fast.cuh"
#include "tendency_tma.cuh"

#ifndef OB_STAGE_LIT
#define OB_STAGE_LIT true   // coefficients as compile-time literals (false: constant-bank tables, for A/B checks)
#endif

namespace ob {

template <typename T, int N, int W, int NCL>
struct StageCfg {
    static constexpr int EPV = 16 / (int)sizeof(T);
    static constexpr int TXC = 32, TYC = W;
    static constexpr int TW = ((TXC + 2 * N + (EPV - 1)) + EPV - 1) / EPV * EPV;   // box width: 16-byte multiple incl. the origin round-down
    static constexpr int TH = TYC + 2 * N;
    static constexpr int BOX_BYTES = TW * TH * (int)sizeof(T);
    static constexpr int PLANE_BYTES = (BOX_BYTES + 127) / 128 * 128;
    static constexpr int PL = PLANE_BYTES / (int)sizeof(T);   // elements between the planes of consecutive fields
    static constexpr int D = N + 2;                           // ring depth: levels k .. k+N + one in flight
    static constexpr int NF = 4;                              // staged fields: u, v, w, tracer of the pass
    static constexpr int NQ = 4;                              // tendencies per pass
    static constexpr int NV = 1 + NCL;                        // flux components: advective + one per closure
    static constexpr int LEVEL_BYTES = NF * PLANE_BYTES;
    static constexpr int LEVEL = LEVEL_BYTES / (int)sizeof(T);
    static constexpr int RING_BYTES = D * LEVEL_BYTES;
    static constexpr int YX_SLOT = (W + 1) * NQ * NV * 32;    // published south fluxes: [w][q][v][lane], w = W: helper (north edge)
    static constexpr int XE_SLOT = NQ * NV * 32;              // published east-edge fluxes: [q][v][row]
    static constexpr int XCH_SLOT = YX_SLOT + XE_SLOT;
    static constexpr int XCH_BYTES = 2 * XCH_SLOT * (int)sizeof(T);
    static constexpr int NH = 4;                              // helper warps: helper q serves tendency q
    static constexpr int NBAR = 2 * D + 2 * (W + 1);
    static constexpr int SMEM_BYTES = RING_BYTES + XCH_BYTES + NBAR * 8 + D * 4;   // + the release counters of the TMA path
    static constexpr int THREADS = (W + NH) * 32;
    static constexpr bool FITS = SMEM_BYTES <= 227 * 1024;
};

// tile height (= compute warps): 16 rows; 12 when two closures add a third published flux component
constexpr int stage_rows(int ncl) { return ncl >= 2 ? 12 : 16; }

// What the four staged fields and the (up to) four tendencies of a CTA are:
//   MT  momentum + one tracer per pass: slots (u, v, w, c_t); tendencies Gu, Gv, Gw (pass 0 only), Gc_t.  Every closure is a
//       ScalarDiffusivity (constants, no closure fields).
//   MN  momentum of an LES model: slots (u, v, w, nu_e); tendencies Gu, Gv, Gw.  Closure 0 is the eddy-viscosity closure
//       (Smagorinsky or AMD) whose nu_e is staged like a velocity component; closure 1, if any, is a ScalarDiffusivity.
//   TT  two tracers of an LES model per pass: slots (c_A, kappa_A, c_B, kappa_B); tendencies Gc_A, Gc_B.  kappa is the AMD
//       kappa_e of the tracer, or nu_e for Smagorinsky (divided by Pr at the face); the advecting velocities are single
//       own-point values and come straight from global memory (three coalesced loads per cell and level).
enum { STAGE_MT = 0, STAGE_MN = 1, STAGE_TT = 2 };

struct StageLaunch {
    int ntx, nty, nkc;      // tiles in x, y and k-chunks
    int kbeg, kend, klen;   // levels kbeg .. kend in chunks of klen
    int use_tma;            // 0: the helper warps copy with cp.async (row pitch not 16-byte aligned)
    int npass;              // gridDim.y
};
#define OB_STAGE_MAXPASS 8
struct StageMaps {
    CUtensorMap m[OB_STAGE_MAXPASS][4];   // tensor map of slot s in pass p
};

__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// wait of a warp that is expected to be early (the helpers run ahead of the compute warps): the try_wait carries a
// suspend-time hint, so the warp sleeps in hardware instead of competing for issue slots with a polling loop
__device__ __forceinline__ void mbar_wait_long(uint64_t *b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(20000u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cp_async_elem(void *dst, const void *src, int bytes, bool valid) {
    const int sz = valid ? bytes : 0;
    if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

// The levels k .. k+N of the ring as seen from one point of the tile: pl[d] points at slot 0 of level k+d, already
// offset to that point; slot f is PL elements further, a row TW elements.
template <typename T, int N, int TW, int PL>
struct StageView {
    const T *pl[N + 1];
    __device__ __forceinline__ T at(int f, int d, int dx, int dy) const { return pl[d][f * PL + dy * TW + dx]; }
};

// own-column value of slot f at level k+dl, dl in [-(N-1), N]: below k from the register history (h[m] = level k-(N-1)+m)
template <typename T, int N, int TW, int PL>
__device__ __forceinline__ T zval(const StageView<T, N, TW, PL> &V, const T (&h)[N - 1], int f, int dl) {
    if (dl < 0) return h[N - 1 + dl];
    return V.at(f, dl, 0, 0);
}

// Advective flux of a tendency of kind WHICH (0 u, 1 v, 2 w, 3 tracer) through the face owned in direction ADV (0: west,
// 1: south, 2: upper face k+1), from the staged planes; the operands and their order are those of fast_flux
// (tendency_fast.cuh).  QS: slot of the advected field.  hq: its history; ha: history of the advecting component whose
// z-line is needed (u for ADV 0, v for 1, w for 2).  GV: the advecting velocity of a tracer flux is passed in `vel`
// (tracer-only passes read it from global memory) instead of being taken from slot ADV.
template <typename T, int N, int WHICH, int ADV, bool STR, int TW, int PL, int QS = WHICH, bool GV = false>
__device__ __forceinline__ T stage_flux(const StageView<T, N, TW, PL> &V, const T (&hq)[N - 1], const T (&ha)[N - 1], const FastGeom<T, STR> &g, int k,
                                        T vel = T(0)) {
    constexpr int NC = N - 1;
    constexpr int AF = ADV;      // slot of the advecting component
    constexpr int LV = ADV == 2 ? 1 : 0;   // the upper face belongs to level k+1
    T s[2 * N];
#pragma unroll
    for (int m = 0; m < 2 * N; m++) {
        if constexpr (ADV == 0) s[m] = V.at(QS, 0, m - N, 0);
        else if constexpr (ADV == 1) s[m] = V.at(QS, 0, 0, m - N);
        else s[m] = zval<T, N, TW, PL>(V, hq, QS, 1 - N + m);
    }
    if constexpr (WHICH == 3) {
        T a[1];
        if constexpr (GV) a[0] = vel;
        else a[0] = V.at(AF, LV, 0, 0);
        return flux_from_values<T, N, true, WHICH, ADV, STR, OB_STAGE_LIT>(s, a, g, k + LV);
    } else {
        T a[2 * NC];
#pragma unroll
        for (int m = 0; m < 2 * NC; m++) {
            if constexpr (WHICH == 0) a[m] = V.at(AF, LV, m - NC, 0);
            else if constexpr (WHICH == 1) a[m] = V.at(AF, LV, 0, m - NC);
            else a[m] = zval<T, N, TW, PL>(V, ha, AF, LV - NC + m);
        }
        return flux_from_values<T, N, true, WHICH, ADV, STR, OB_STAGE_LIT>(s, a, g, k + LV);
    }
}

// Non-advective terms from the staged planes: FastTerms (tendency_fast.cuh) with every load redirected to shared memory
// (level offset -1: register history of the own column; 0, +1: ring).  MODE / NCL / KL fix at compile time which closures
// exist and where their fields are, so no closure-kind branch and no global closure-field load is left in the code:
//   closure 0 of an LES mode (KL = CL_SMAG or CL_AMD): nu_e in slot 3 (MN), kappa in slot CS+1 (TT); every other closure:
//   ScalarDiffusivity constants.  pHY' and a non-staged buoyancy tracer stay on the global path.
template <typename T, int N, int MODE, int NCL, int KL, bool STR, int TW, int PL>
struct StageTerms {
    const TendP<T> &P;
    const FastGeom<T, STR> &G;
    const StageView<T, N, TW, PL> &V;
    const T (&h0)[N - 1], (&h1)[N - 1], (&h2)[N - 1], (&h3)[N - 1];   // own-column histories of the four slots
    int eo;       // global element offset of the (clamped) point at level k
    int k;
    int tr;       // tracer index of the tracer tendency (MT: the pass's tracer; TT: set per tracer)
    int cs;       // TT: slot of that tracer (its kappa is in slot cs + 1)
    T ixp, iyp;   // MN: If1x / If1y of nu_e at this point one level below (carried in registers)
    static constexpr bool LES0 = MODE != STAGE_MT;
    __device__ __forceinline__ T slot(int f, int a, int b, int c) const {
        if (c < 0) return f == 0 ? h0[N - 2] : f == 1 ? h1[N - 2] : f == 2 ? h2[N - 2] : h3[N - 2];
        return V.at(f, c, a, b);
    }
    __device__ __forceinline__ T ldg(const T *p, int a, int b, int c) const { return __ldg(p + (a + b * G.sy + c * G.sz)); }
    __device__ __forceinline__ const T *at(const Fld<T> &f) const { return f.p + f.off + eo; }
    __device__ __forceinline__ T dzC(int c) const { return G.dzC(k + c); }
    __device__ __forceinline__ T dzF(int c) const { return G.dzF(k + c); }
    __device__ __forceinline__ T dx_u(int a, int b, int c) const { return (slot(0, a + 1, b, c) - slot(0, a, b, c)) * G.rdx; }
    __device__ __forceinline__ T dy_v(int a, int b, int c) const { return (slot(1, a, b + 1, c) - slot(1, a, b, c)) * G.rdy; }
    __device__ __forceinline__ T dz_w(int a, int b, int c) const { return (slot(2, a, b, c + 1) - slot(2, a, b, c)) * G.rdzC(k + c); }
    __device__ __forceinline__ T dx_v(int a, int b, int c) const { return (slot(1, a, b, c) - slot(1, a - 1, b, c)) * G.rdx; }
    __device__ __forceinline__ T dy_u(int a, int b, int c) const { return (slot(0, a, b, c) - slot(0, a, b - 1, c)) * G.rdy; }
    __device__ __forceinline__ T dx_w(int a, int b, int c) const { return (slot(2, a, b, c) - slot(2, a - 1, b, c)) * G.rdx; }
    __device__ __forceinline__ T dz_u(int a, int b, int c) const { return (slot(0, a, b, c) - slot(0, a, b, c - 1)) * G.rdzF(k + c); }
    __device__ __forceinline__ T dy_w(int a, int b, int c) const { return (slot(2, a, b, c) - slot(2, a, b - 1, c)) * G.rdy; }
    __device__ __forceinline__ T dz_v(int a, int b, int c) const { return (slot(1, a, b, c) - slot(1, a, b, c - 1)) * G.rdzF(k + c); }
    __device__ __forceinline__ T S12(int a, int b, int c) const { return T(0.5) * add_rn(dy_u(a, b, c), dx_v(a, b, c)); }
    __device__ __forceinline__ T S13(int a, int b, int c) const { return T(0.5) * add_rn(dz_u(a, b, c), dx_w(a, b, c)); }
    __device__ __forceinline__ T S23(int a, int b, int c) const { return T(0.5) * add_rn(dz_v(a, b, c), dy_w(a, b, c)); }
    // two-point interpolation of the ccc array in slot f to the face in direction D (interpolation_operators.jl:8-28); level -1
    // of the x / y interpolants at the own point comes from the carried registers
    __device__ __forceinline__ T If1(int f, int D, int a, int b, int c) const {
        if (c < 0) return D == 0 ? ixp : iyp;
        return T(0.5) * (V.at(f, c - (D == 2), a - (D == 0), b - (D == 1)) + V.at(f, c, a, b));
    }
    __device__ __forceinline__ T If2(int f, int D2, int D1, int a, int b, int c) const {
        return T(0.5) * (If1(f, D1, a - (D2 == 0), b - (D2 == 1), c - (D2 == 2)) + If1(f, D1, a, b, c));
    }
    // viscosity of closure m at ccc / ffc / fcf / cff (abstract_scalar_diffusivity_closure.jl:330-351)
    __device__ __forceinline__ T nu_ccc(int m, int a, int b, int c) const { if (LES0 && m == 0) return V.at(3, c, a, b); return P.cl[m].nu; }
    __device__ __forceinline__ T nu_ffc(int m, int a, int b, int c) const { if (LES0 && m == 0) return If2(3, 1, 0, a, b, c); return P.cl[m].nu; }
    __device__ __forceinline__ T nu_fcf(int m, int a, int b, int c) const { if (LES0 && m == 0) return If2(3, 2, 0, a, b, c); return P.cl[m].nu; }
    __device__ __forceinline__ T nu_cff(int m, int a, int b, int c) const { if (LES0 && m == 0) return If2(3, 2, 1, a, b, c); return P.cl[m].nu; }
    __device__ __forceinline__ T ux(int m, int a, int b, int c) const { return (G.dy * dzC(c)) * (-2 * (nu_ccc(m, a, b, c) * dx_u(a, b, c))); }
    __device__ __forceinline__ T uy(int m, int a, int b, int c) const { return (G.dx * dzC(c)) * (-2 * (nu_ffc(m, a, b, c) * S12(a, b, c))); }
    __device__ __forceinline__ T uz(int m, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_fcf(m, a, b, c) * S13(a, b, c))); }
    __device__ __forceinline__ T vx(int m, int a, int b, int c) const { return (G.dy * dzC(c)) * (-2 * (nu_ffc(m, a, b, c) * S12(a, b, c))); }
    __device__ __forceinline__ T vy(int m, int a, int b, int c) const { return (G.dx * dzC(c)) * (-2 * (nu_ccc(m, a, b, c) * dy_v(a, b, c))); }
    __device__ __forceinline__ T vz(int m, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_cff(m, a, b, c) * S23(a, b, c))); }
    __device__ __forceinline__ T wx(int m, int a, int b, int c) const { return (G.dy * dzF(c)) * (-2 * (nu_fcf(m, a, b, c) * S13(a, b, c))); }
    __device__ __forceinline__ T wy(int m, int a, int b, int c) const { return (G.dx * dzF(c)) * (-2 * (nu_cff(m, a, b, c) * S23(a, b, c))); }
    __device__ __forceinline__ T wz(int m, int a, int b, int c) const { return (G.dx * G.dy) * (-2 * (nu_ccc(m, a, b, c) * dz_w(a, b, c))); }
    // diffusive flux of the tracer along D at the face (a, b, c) (abstract_scalar_diffusivity_closure.jl:260-262)
    __device__ __forceinline__ T qflux(int m, int D, int a, int b, int c) const {
        const int fc = MODE == STAGE_TT ? cs : 3;   // slot of the tracer
        T kap;
        if (LES0 && m == 0) {
            if constexpr (KL == CL_SMAG) kap = If1(fc + 1, D, a, b, c) / P.cl[0].Pr[tr];
            else kap = If1(fc + 1, D, a, b, c);
        } else {
            kap = P.cl[m].kappa[tr];
        }
        const T rd = D == 0 ? G.rdx : D == 1 ? G.rdy : G.rdzF(k + c);
        const T A = D == 0 ? G.dy * dzC(c) : D == 1 ? G.dx * dzC(c) : G.dx * G.dy;
        const T dc = (slot(fc, a, b, c) - slot(fc, a - (D == 0), b - (D == 1), c - (D == 2))) * rd;
        return A * (-kap * dc);
    }
    // closure flux of a tendency of kind WHICH through the owned face in direction D (D == 2: the UPPER face)
    template <int WHICH, int D> __device__ __forceinline__ T own_closure_flux(int m) const {
        if constexpr (WHICH == 3) return qflux(m, D, 0, 0, D == 2 ? 1 : 0);
        else if constexpr (WHICH == 0) return D == 0 ? ux(m, -1, 0, 0) : D == 1 ? uy(m, 0, 0, 0) : uz(m, 0, 0, 1);
        else if constexpr (WHICH == 1) return D == 0 ? vx(m, 0, 0, 0) : D == 1 ? vy(m, 0, -1, 0) : vz(m, 0, 0, 1);
        else return D == 0 ? wx(m, 0, 0, 0) : D == 1 ? wy(m, 0, 0, 0) : wz(m, 0, 0, 0);
    }
    __device__ __forceinline__ T bpert(int c) const {
        if (P.buoy == BUOY_TRACER) {
            if (MODE == STAGE_MT && P.ib == tr) return slot(3, 0, 0, c);
            return ldg(at(P.c[P.ib]), 0, 0, c);
        }
        if (P.buoy == BUOY_SEAWATER) return P.grav * (P.alpha * ldg(at(P.c[P.iT]), 0, 0, c) - P.beta * ldg(at(P.c[P.iS]), 0, 0, c));
        return 0;
    }
    // the tendency assemblers, same term order as FastTerms::finish<WHICH, true>
    template <int WHICH> __device__ __forceinline__ T finish(T adv, T closure_term) const {
        T r = -adv;
        if constexpr (WHICH == 0) {
            if (P.has_cor) {
                const T fbar = T(0.5) * (P.f + P.f);
                const T A = G.dx * dzC(0);
                const T I = interp4_rn(A, slot(1, -1, 0, 0), slot(1, 0, 0, 0), slot(1, -1, 1, 0), slot(1, 0, 1, 0));
                r = sub_rn(r, mul_rn(mul_rn(-fbar, I), 1 / (G.dx * dzC(0))));
            }
            if (P.has_pHY) { const T *ph = at(P.pHY); r = sub_rn(r, mul_rn(ldg(ph, 0, 0, 0) - ldg(ph, -1, 0, 0), G.rdx)); }
        } else if constexpr (WHICH == 1) {
            if (P.has_cor) {
                const T fbar = T(0.5) * (P.f + P.f);
                const T A = G.dy * dzC(0);
                const T I = interp4_rn(A, slot(0, 0, -1, 0), slot(0, 1, -1, 0), slot(0, 0, 0, 0), slot(0, 1, 0, 0));
                r = sub_rn(r, mul_rn(mul_rn(fbar, I), 1 / (G.dy * dzC(0))));
            }
            if (P.has_pHY) { const T *ph = at(P.pHY); r = sub_rn(r, mul_rn(ldg(ph, 0, 0, 0) - ldg(ph, 0, -1, 0), G.rdy)); }
        } else if constexpr (WHICH == 2) {
            if (!P.has_pHY && P.buoy != BUOY_NONE) r = r + T(0.5) * (bpert(-1) + bpert(0));
        }
        if constexpr (NCL > 0) r = sub_rn(r, closure_term);
        return r;
    }
};

template <typename T, int N, int W, int MODE, int NCL, int KL, bool STR>
__global__ void __launch_bounds__((W + 4) * 32, 1) tendency_stage_kernel(const __grid_constant__ TendP<T> P, const __grid_constant__ StageMaps M, const StageLaunch L) {
    using C = StageCfg<T, N, W, NCL>;
    constexpr int TW = C::TW, PL = C::PL, D = C::D, NV = C::NV;
    constexpr int NCLR = NCL > 0 ? NCL : 1;
    using View = StageView<T, N, TW, PL>;
    using Terms = StageTerms<T, N, MODE, NCL, KL, STR, TW, PL>;
    extern __shared__ __align__(128) unsigned char smem[];
    T *ring = reinterpret_cast<T *>(smem);
    T *xch = reinterpret_cast<T *>(smem + C::RING_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::RING_BYTES + C::XCH_BYTES);
    uint64_t *full = bars, *empty = bars + D, *pub = bars + 2 * D;   // pub[2*w + slot], w = 1 .. W (W: the helpers)
    int *done = reinterpret_cast<int *>(smem + C::RING_BYTES + C::XCH_BYTES + C::NBAR * 8);   // TMA path: releases per slot (monotonic)

    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps role branches and the
    // coefficient literals on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0), lane = (int)threadIdx.x & 31;
    const GridD<T> &gg = P.g;
    const int Nx = gg.N[0], Ny = gg.N[1];
    int b = blockIdx.x;
    const int tile_x = b % L.ntx; b /= L.ntx;
    const int tile_y = b % L.nty;
    const int kc = b / L.nty;
    const int pass = blockIdx.y;
    const int i0 = 1 + tile_x * C::TXC, j0 = 1 + tile_y * C::TYC;
    const int k0 = L.kbeg + kc * L.klen, k1 = min(k0 + L.klen - 1, L.kend);
    const int kfirst = k0 - N, klast = k1 + N;

    // ---- what this pass stages and computes (CTA-uniform) ----------------------------------------------------------------
    // tendency slot q: active?  MT: q = 0..2 momentum (pass 0), q = 3 tracer `pass`;  MN: q = 0..2;  TT: q = 0, 1 tracers 2*pass, 2*pass+1
    const int trA = MODE == STAGE_MT ? pass : 2 * pass, trB = 2 * pass + 1;
    const bool a012 = MODE == STAGE_MT ? pass == 0 : true;
    const bool act[4] = {MODE == STAGE_TT ? true : a012, MODE == STAGE_TT ? trB < P.ntr : a012, MODE == STAGE_TT ? false : a012, MODE == STAGE_MT ? trA < P.ntr : false};
    const T *slotp[4];   // parent array of each slot (cp.async path, and which slots exist)
    if constexpr (MODE == STAGE_MT) {
        slotp[0] = P.u.p; slotp[1] = P.v.p; slotp[2] = P.w.p; slotp[3] = act[3] ? P.c[trA].p : nullptr;
    } else if constexpr (MODE == STAGE_MN) {
        slotp[0] = P.u.p; slotp[1] = P.v.p; slotp[2] = P.w.p; slotp[3] = P.nue[0].p;
    } else {
        slotp[0] = P.c[trA].p; slotp[1] = KL == CL_AMD ? P.kappae[0][trA].p : P.nue[0].p;
        slotp[2] = act[1] ? P.c[trB].p : nullptr; slotp[3] = act[1] ? (KL == CL_AMD ? P.kappae[0][trB].p : P.nue[0].p) : nullptr;
    }
    const bool have[4] = {true, true, MODE != STAGE_TT || act[1], MODE == STAGE_MT ? act[3] : MODE == STAGE_MN ? true : act[1]};   // which slots are staged
    const int nfields = (int)have[0] + (int)have[1] + (int)have[2] + (int)have[3];

    // tile origin in parent coordinates (0-based); the box starts on a 16-byte boundary of the row
    const int cxu = i0 - N + gg.H[0] - 1;
    const int cx0 = cxu & ~(C::EPV - 1), cy0 = j0 - N + gg.H[1] - 1;
    const int sh = cxu - cx0;
    const bool use_tma = L.use_tma != 0;
    auto tma_issue = [&](int Lv) {   // one thread: the boxes of level Lv into its ring slot
        const int s = (Lv - kfirst) % D;
        unsigned char *dst = smem + s * C::LEVEL_BYTES;
        mbar_expect_tx(&full[s], nfields * C::BOX_BYTES);
        const int cz = Lv + gg.H[2] - 1;
        // (static indices into the __grid_constant__ maps: a run-time index would make the compiler copy them to local memory)
#pragma unroll
        for (int p = 0; p < OB_STAGE_MAXPASS; p++) {
            if (p != pass) continue;
#pragma unroll
            for (int f = 0; f < 4; f++)
                if (have[f]) tma_load_3d(dst + f * C::PLANE_BYTES, &M.m[p][f], &full[s], cx0, cy0, cz);
        }
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < D; s++) { mbar_init(&full[s], use_tma ? 1 : nfields * 32); mbar_init(&empty[s], W + C::NH); done[s] = 0; }
        for (int s = 0; s < 2 * (W + 1); s++) mbar_init(&pub[s], s >= 2 * W ? C::NH : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (use_tma)
            for (int Lv = kfirst; Lv < kfirst + D && Lv <= klast; Lv++) tma_issue(Lv);
    }
    __syncthreads();

    // ---------------------------------------------------------------- consumers ------------------------------------
    // Compute warps (warp < W): own point (i0+lane, j0+warp), the same for every phase.  Helpers (warp >= W): the south-flux
    // point is (i0+lane, j0+TYC) -- the tile's north edge, lanes = columns -- and the west-flux point (i0+32, j0+lane) --
    // the east edge, lanes = rows; helper hq does tendency slot hq.  Phase A (south fluxes) is the same code in every warp.
    const bool helper = warp >= W;
    const int hq = warp - W;
    const bool do0 = act[0] && (!helper || hq == 0), do1 = act[1] && (!helper || hq == 1), do2 = act[2] && (!helper || hq == 2);
    const bool do3 = act[3] && (!helper || hq == 3);
    const int pw = helper ? W : warp;   // publication row of this warp's south fluxes
    FastGeom<T, STR> g;
    g.init(gg, P.u.sy, P.u.sz);
    const int rowB = min(lane, C::TYC - 1);
    const int ownY = (helper ? C::TYC + N : warp + N) * TW + lane + N + sh;
    const int ownX = helper ? (rowB + N) * TW + 32 + N + sh : ownY;
    const int gi = i0 + lane, gj = j0 + (helper ? 0 : warp);
    const int eoY = min(gi, Nx + 1) + min(helper ? j0 + C::TYC : gj, Ny + 1) * g.sy;
    const int eoX = helper ? min(i0 + 32, Nx + 1) + min(j0 + rowB, Ny + 1) * g.sy : eoY;
    const bool live = !helper && gi <= Nx && gj <= Ny;
    auto yx_at = [&](int slot, int w, int q, int v) -> T * { return xch + slot * C::XCH_SLOT + ((w * C::NQ + q) * NV + v) * 32 + lane; };
    auto xe_at = [&](int slot, int q, int v, int row) -> T * { return xch + slot * C::XCH_SLOT + C::YX_SLOT + (q * NV + v) * 32 + row; };

    // register state carried from level to level: own-column history of every slot (helpers, MT / MN: h[0] = u at the west-flux
    // point, h[1] = v at the south-flux point), lower-face fluxes, and (MN) the x / y face interpolants of nu_e one level below
    T h[4][N - 1], lower[4], lower_c[4][NCLR];
    T ixp = T(0), iyp = T(0);
    T gun = T(0), gvn = T(0), gwn = T(0);   // TT: advecting velocities of the next level
#pragma unroll
    for (int q = 0; q < 4; q++) {
        lower[q] = T(0);
#pragma unroll
        for (int m = 0; m < NCLR; m++) lower_c[q][m] = T(0);
#pragma unroll
        for (int m = 0; m < N - 1; m++) h[q][m] = T(0);
    }
    // running ring positions: level k is in slot sk; level k+N (the newest one this level needs) in slot sn with phase pn
    int sk = 0, sn = N % D, pn = (N / D) & 1;
    // cp.async path: the 32 lanes of helper f copy the box of slot f element by element (zero-fill outside the parent
    // array) for every level up to `upto`; a ring slot is reused once every consumer has released its previous level.
    int fed = kfirst - 1;
    auto feed = [&](bool blocking, int upto) {
        const int Px = P.u.sy, Py = (int)(P.u.sz / P.u.sy);
        const T *base;
        if constexpr (MODE == STAGE_MT) base = (hq == 0 ? P.u : hq == 1 ? P.v : hq == 2 ? P.w : P.c[trA]).p;
        else base = hq == 0 ? slotp[0] : hq == 1 ? slotp[1] : hq == 2 ? slotp[2] : slotp[3];
        while (fed < upto && fed < klast) {
            const int Lv = fed + 1, n = Lv - kfirst, s = n % D;
            if (n >= D) {
                const uint32_t par = (uint32_t)(((n / D) - 1) & 1);
                if (blocking) mbar_wait_long(&empty[s], par);
                else {
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&empty[s])), "r"(par) : "memory");
                    if (!ok) return;
                }
            }
            T *dst = reinterpret_cast<T *>(smem + s * C::LEVEL_BYTES + hq * C::PLANE_BYTES);
            const T *src = base + (long)(Lv + gg.H[2] - 1) * P.u.sz;
            for (int e = lane; e < C::TW * C::TH; e += 32) {
                const int yy = e / C::TW, xx = e - yy * C::TW;
                const int gx = cx0 + xx, gy = cy0 + yy;
                const bool ok = gx < Px && gy < Py;
                cp_async_elem(dst + e, src + (ok ? (long)gy * Px + gx : 0), (int)sizeof(T), ok);
            }
            cp_async_mbar_arrive(&full[s]);
            fed = Lv;
        }
    };
    const bool feeder = !use_tma && helper && (hq == 0 ? have[0] : hq == 1 ? have[1] : hq == 2 ? have[2] : have[3]);
    if (feeder) feed(true, kfirst + N - 1);
    for (int n = 0; n < N; n++) mbar_wait(&full[n], 0);

    for (int k = kfirst; k <= k1; k++) {
        if (feeder) feed(true, k + N);
        // The helpers have a sixth of a compute warp's work: left to themselves they would spend the level polling the ring
        // barrier and take issue slots from the compute warps.  They block on a hardware named barrier instead, which
        // compute warp 0 arrives at (without waiting) once it has seen level k+N in the ring.
        // Two barrier ids alternate with the level: warp 0 can be up to two levels ahead of a helper while the ring is still
        // being primed (afterwards a level k+N is only loaded once every consumer has released k-2), and arrivals of
        // different levels must not meet in one barrier phase.
        const int bar_id = 1 + ((k - kfirst) & 1);
        if (helper) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * (1 + C::NH)) : "memory");
        mbar_wait(&full[sn], (uint32_t)pn);
        if (warp == 0) asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "n"(32 * (1 + C::NH)) : "memory");
        const T *lev[N + 1];
#pragma unroll
        for (int d = 0; d <= N; d++) { int s = sk + d; if (s >= D) s -= D; lev[d] = ring + s * C::LEVEL; }
        const int e = k - k0, xs = e & 1, xp = (e >> 1) & 1;
        const bool full_level = k >= k0;
        View VY;
#pragma unroll
        for (int d = 0; d <= N; d++) VY.pl[d] = lev[d] + ownY;
        Terms FY{P, g, VY, h[0], h[1], h[2], h[3], eoY + k * g.sz, k, trA, 0, ixp, iyp};
        // TT: the advecting velocities at the own points, from global memory (u at the west-flux point, v at the south-flux
        // point, w one level up)
        // (loaded one level ahead: the three loads of level k+1 are in flight while level k is computed)
        T gu = T(0), gv = T(0), gw = T(0);
        if constexpr (MODE == STAGE_TT) {
            gu = gun; gv = gvn; gw = gwn;
            if (k + 1 >= k0 - 1 && k + 1 <= k1) {
                gun = __ldg(P.u.p + P.u.off + eoX + (k + 1) * g.sz);
                gvn = __ldg(P.v.p + P.v.off + eoY + (k + 1) * g.sz);
                gwn = __ldg(P.w.p + P.w.off + eoY + (k + 2) * g.sz);
            }
        }
        // one tendency slot q: kind of tendency, slot of its field, tracer index
        auto south = [&](auto qtag) {   // phase A of slot q
            constexpr int Q = decltype(qtag)::value;
            constexpr int WHICH = MODE == STAGE_TT ? 3 : Q;
            constexpr int QS = MODE == STAGE_TT ? 2 * Q : Q;
            if constexpr (MODE == STAGE_TT) { FY.tr = Q == 0 ? trA : trB; FY.cs = QS; }
            *yx_at(xs, pw, Q, 0) = stage_flux<T, N, WHICH, 1, STR, TW, PL, QS, MODE == STAGE_TT>(VY, h[QS], h[1], g, k, gv);
#pragma unroll
            for (int m = 0; m < NCL; m++) *yx_at(xs, pw, Q, 1 + m) = FY.template own_closure_flux<WHICH, 1>(m);
        };
        if (full_level) {
            // ---- phase A: south fluxes, published for the warp below (helpers: the tile's north edge) ------------------
            if (do0) south(std::integral_constant<int, 0>{});
            if (do1) south(std::integral_constant<int, 1>{});
            if constexpr (MODE != STAGE_TT) {
                if (do2) south(std::integral_constant<int, 2>{});
                if constexpr (MODE == STAGE_MT) { if (do3) south(std::integral_constant<int, 3>{}); }
            }
            if (helper) {
                // ---- helper phase B: west fluxes at the tile's east edge, lanes = rows -----------------------------------
                View VX;
#pragma unroll
                for (int d = 0; d <= N; d++) VX.pl[d] = lev[d] + ownX;
                Terms FX{P, g, VX, h[0], h[1], h[2], h[3], eoX + k * g.sz, k, trA, 0, ixp, iyp};
                auto west_edge = [&](auto qtag) {
                    constexpr int Q = decltype(qtag)::value;
                    constexpr int WHICH = MODE == STAGE_TT ? 3 : Q;
                    constexpr int QS = MODE == STAGE_TT ? 2 * Q : Q;
                    if constexpr (MODE == STAGE_TT) { FX.tr = Q == 0 ? trA : trB; FX.cs = QS; }
                    *xe_at(xs, Q, 0, lane) = stage_flux<T, N, WHICH, 0, STR, TW, PL, QS, MODE == STAGE_TT>(VX, h[QS], h[0], g, k, gu);
#pragma unroll
                    for (int m = 0; m < NCL; m++) *xe_at(xs, Q, 1 + m, lane) = FX.template own_closure_flux<WHICH, 0>(m);
                };
                if (lane < C::TYC) {
                    if (do0) west_edge(std::integral_constant<int, 0>{});
                    if (do1) west_edge(std::integral_constant<int, 1>{});
                    if constexpr (MODE != STAGE_TT) {
                        if (do2) west_edge(std::integral_constant<int, 2>{});
                        if constexpr (MODE == STAGE_MT) { if (do3) west_edge(std::integral_constant<int, 3>{}); }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&pub[2 * W + xs]);
            } else {
                __syncwarp();
                if (lane == 0 && warp > 0) mbar_arrive(&pub[2 * warp + xs]);
            }
        }
        if (!helper && k >= k0 - 1) {
            // ---- phase C: upper fluxes; on full levels the west fluxes, the divergence and the tendency ----------------
            const int eo = eoY + k * g.sz;
            bool waited = false;   // the published fluxes are awaited as late as possible: right before their first use
            auto tendency = [&](auto qtag) {
                constexpr int Q = decltype(qtag)::value;
                constexpr int WHICH = MODE == STAGE_TT ? 3 : Q;
                constexpr int QS = MODE == STAGE_TT ? 2 * Q : Q;
                if constexpr (MODE == STAGE_TT) { FY.tr = Q == 0 ? trA : trB; FY.cs = QS; }
                const Fld<T> &G = MODE == STAGE_TT ? P.Gc[Q == 0 ? trA : trB] : Q == 0 ? P.Gu : Q == 1 ? P.Gv : Q == 2 ? P.Gw : P.Gc[trA];
                const T upper = stage_flux<T, N, WHICH, 2, STR, TW, PL, QS, MODE == STAGE_TT>(VY, h[QS], h[2], g, k, gw);
                T cup[NCLR];
#pragma unroll
                for (int m = 0; m < NCL; m++) cup[m] = FY.template own_closure_flux<WHICH, 2>(m);
                if (full_level) {
                    const T fx = stage_flux<T, N, WHICH, 0, STR, TW, PL, QS, MODE == STAGE_TT>(VY, h[QS], h[0], g, k, gu);
                    if (!waited) {
                        mbar_wait(&pub[2 * (warp + 1) + xs], (uint32_t)xp);
                        if (warp + 1 != W) mbar_wait(&pub[2 * W + xs], (uint32_t)xp);
                        waited = true;
                    }
                    const T fy = *yx_at(xs, warp, Q, 0);
                    const T fy1 = *yx_at(xs, warp + 1, Q, 0);
                    T fx1 = __shfl_down_sync(0xffffffffu, fx, 1);
                    if (lane == 31) fx1 = *xe_at(xs, Q, 0, warp);
                    const T Vi = WHICH == 2 ? g.rVf(k) : g.rVc(k);
                    const T adv = Vi * ((fx1 - fx) + (fy1 - fy) + (upper - lower[Q]));
                    T term = T(0);
#pragma unroll
                    for (int m = 0; m < NCL; m++) {
                        const T cx = FY.template own_closure_flux<WHICH, 0>(m);
                        const T cyv = *yx_at(xs, warp, Q, 1 + m);
                        const T cy1 = *yx_at(xs, warp + 1, Q, 1 + m);
                        T cx1 = __shfl_down_sync(0xffffffffu, cx, 1);
                        if (lane == 31) cx1 = *xe_at(xs, Q, 1 + m, warp);
                        // (the marching kernel forms these under run-time closure counts, where nothing contracts; pinned here)
                        const T d = mul_rn(Vi, (cx1 - cx) + (cy1 - cyv) + (cup[m] - lower_c[Q][m]));
                        term = m == 0 ? d : add_rn(term, d);
                    }
                    const T res = FY.template finish<WHICH>(adv, term);
                    if (live) G.p[G.off + eo] = res;
                }
                lower[Q] = upper;
#pragma unroll
                for (int m = 0; m < NCL; m++) lower_c[Q][m] = cup[m];
            };
            if (act[0]) tendency(std::integral_constant<int, 0>{});
            if (act[1]) tendency(std::integral_constant<int, 1>{});
            if constexpr (MODE != STAGE_TT) {
                if (act[2]) tendency(std::integral_constant<int, 2>{});
                if constexpr (MODE == STAGE_MT) { if (act[3]) tendency(std::integral_constant<int, 3>{}); }
            }
        }
        // level k becomes k-1: history shift (helpers, MT / MN: u at the west-flux point, v at the south-flux point) and (MN) the
        // face interpolants of nu_e that next level's wx / wy need
        if constexpr (MODE == STAGE_MN) {
            ixp = T(0.5) * (lev[0][3 * PL + ownX - 1] + lev[0][3 * PL + ownX]);
            iyp = T(0.5) * (lev[0][3 * PL + ownY - TW] + lev[0][3 * PL + ownY]);
        }
#pragma unroll
        for (int f = 0; f < 4; f++) {
#pragma unroll
            for (int m = 0; m + 1 < N - 1; m++) h[f][m] = h[f][m + 1];
            h[f][N - 2] = lev[0][f * PL + ((f == 0 && MODE != STAGE_TT) ? ownX : ownY)];
        }
        __syncwarp();
        if (use_tma) {
            // release level k; the last of the W+NH consumer warps to do so issues the loads of level k+D into the slot
            if (lane == 0) {
                __threadfence_block();
                const int old = atomicAdd(&done[sk], 1);
                if (old % (W + C::NH) == W + C::NH - 1 && k + D <= klast) {
                    __threadfence_block();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tma_issue(k + D);
                }
            }
        } else {
            if (lane == 0) mbar_arrive(&empty[sk]);
            // cp.async path: helper f copies slot f.  Opportunistic: if the level that takes this ring slot can be loaded already,
            // do it now; otherwise the blocking catch-up at the top of a later level does it.
            if (feeder) feed(false, k + N + 1);
        }
        if (++sk == D) sk = 0;
        if (++sn == D) { sn = 0; pn ^= 1; }
    }
}

}  // namespace ob
