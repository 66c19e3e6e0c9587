// stage_inst.cu -- one instantiation of the staged-ring tendency kernel (regular and stretched z); compiled once per
// (float type, buffer, mode, closure count, eddy-viscosity kind) with -DOB_SI_T=double -DOB_SI_TN=f64 -DOB_SI_N=3 -DOB_SI_MODE=0
// -DOB_SI_NCL=1 -DOB_SI_KL=0 (see oceananigans.jl_b200/build.py).
#include <string.h>
#include "tendency_stage.cuh"
#include "tma_maps.h"
#include "stage_launch.h"

namespace ob {

template <typename T, int N, int W, int MODE, int NCL, int KL>
static cudaError_t launch_stage_t(const TendP<T> &P, cudaStream_t st, int sm_count, int *nlaunch) {
    using C = StageCfg<T, N, W, NCL>;
    static_assert(C::FITS, "staged-ring tile does not fit in shared memory");
    const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
    StageLaunch L;
    L.ntx = (Nx + C::TXC - 1) / C::TXC;
    L.nty = (Ny + C::TYC - 1) / C::TYC;
    const bool walls = P.g.topo[2] == BOUNDED;
    L.kbeg = walls ? N + 1 : 1;
    L.kend = walls ? Nz - N : Nz;
    L.npass = MODE == STAGE_MT ? (P.ntr > 1 ? P.ntr : 1) : MODE == STAGE_MN ? 1 : (P.ntr + 1) / 2;
    if (L.npass > OB_STAGE_MAXPASS || L.npass < 1) return cudaErrorInvalidValue;
    // chunk length: balance whole waves of CTAs (one CTA per SM) against the N+1 warm-up levels of every chunk
    const long tiles = (long)L.ntx * L.nty * (MODE == STAGE_TT ? L.npass : 1);   // (MT: pass 0 dominates, the tracer passes fill in)
    const int nlev = L.kend - L.kbeg + 1;
    const int sms = sm_count > 0 ? sm_count : 148;
    double best = 1e300;
    int best_nk = 1;
    for (int nk = 1; nk <= nlev; nk++) {
        const int len = (nlev + nk - 1) / nk;
        if (len < 4 && nk > 1) break;
        const long ctas = tiles * nk;
        const long waves = (ctas + sms - 1) / sms;
        const double cost = (double)waves * (len + 2.0);
        if (cost < best) { best = cost; best_nk = nk; }
    }
    L.klen = (nlev + best_nk - 1) / best_nk;
    L.nkc = (nlev + L.klen - 1) / L.klen;
    // slot -> parent array, per pass (the kernel derives the same table for the cp.async path)
    StageMaps M;
    memset(&M, 0, sizeof(M));
    const cuuint64_t Px = (cuuint64_t)P.u.sy, Py = (cuuint64_t)(P.u.sz / P.u.sy);
    L.use_tma = ((Px * sizeof(T)) % 16 == 0 && !getenv("OB_STAGE_NO_TMA")) ? 1 : 0;
    if (L.use_tma) {
        const cuuint64_t Pzc = (cuuint64_t)(Nz + 2 * P.g.H[2]), Pzw = Pzc + (walls ? 1 : 0);
        auto map = [&](int pass, int slot, const T *ptr, bool is_w) {
            return ptr == nullptr || stage_map(&M.m[pass][slot], ptr, Px, Py, is_w ? Pzw : Pzc, C::TW, C::TH, (int)sizeof(T));
        };
        bool ok = true;
        for (int p = 0; ok && p < L.npass; p++) {
            if (MODE == STAGE_MT) {
                ok = map(p, 0, P.u.p, false) && map(p, 1, P.v.p, false) && map(p, 2, P.w.p, true) && map(p, 3, p < P.ntr ? P.c[p].p : nullptr, false);
            } else if (MODE == STAGE_MN) {
                ok = map(p, 0, P.u.p, false) && map(p, 1, P.v.p, false) && map(p, 2, P.w.p, true) && map(p, 3, P.nue[0].p, false);
            } else {
                const int a = 2 * p, b = 2 * p + 1;
                const T *ka = KL == CL_AMD ? P.kappae[0][a].p : P.nue[0].p;
                const T *kb = b < P.ntr ? (KL == CL_AMD ? P.kappae[0][b].p : P.nue[0].p) : nullptr;
                ok = map(p, 0, P.c[a].p, false) && map(p, 1, ka, false) && map(p, 2, b < P.ntr ? P.c[b].p : nullptr, false) && map(p, 3, kb, false);
            }
        }
        if (!ok) L.use_tma = 0;
    }
    if ((long)L.ntx * L.nty * L.nkc > 2147483647L) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)(L.ntx * L.nty * L.nkc), (unsigned)L.npass), block(C::THREADS);
    auto kern = P.g.dzc ? tendency_stage_kernel<T, N, W, MODE, NCL, KL, true> : tendency_stage_kernel<T, N, W, MODE, NCL, KL, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<grid, block, C::SMEM_BYTES, st>>>(P, M, L);
    *nlaunch += 1;
    return cudaGetLastError();
}

#define OB_CAT4_(tn, n, m, c, k) launch_stage_##tn##_n##n##_m##m##_c##c##_k##k
#define OB_CAT4(tn, n, m, c, k) OB_CAT4_(tn, n, m, c, k)

cudaError_t OB_CAT4(OB_SI_TN, OB_SI_N, OB_SI_MODE, OB_SI_NCL, OB_SI_KL)(const TendP<OB_SI_T> &P, cudaStream_t st, int sm_count, int *nlaunch) {
    return launch_stage_t<OB_SI_T, OB_SI_N, stage_rows(OB_SI_NCL), OB_SI_MODE, OB_SI_NCL, OB_SI_KL>(P, st, sm_count, nlaunch);
}

}  // namespace ob
