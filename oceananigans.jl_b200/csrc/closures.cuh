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

// AnisotropicMinimumDissipation: νₑ and every κₑ of one cell in one thread.  The 27 normalised velocity gradients the
// "30 terms" are built from (3 at ccc; ∂x v, ∂y u at the 4 surrounding ffc points; ∂x w, ∂z u at the 4 fcf points;
// ∂y w, ∂z v at the 4 cff points) are evaluated ONCE and every term of anisotropic_minimum_dissipation.jl:251-358 is then
// formed from them with the reference's own expression order.  A Flat direction needs no special case: its index stride
// is 0 (so δ = 0 exactly and ℑ = 0.5*(f+f) = f exactly) and its spacing is 1.
template <typename T>
__device__ __forceinline__ T interp4(const T (&f)[2][2]) {  // ℑ_outer(ℑ_inner f), f[inner][outer]
    return T(0.5) * (T(0.5) * (f[0][0] + f[1][0]) + T(0.5) * (f[0][1] + f[1][1]));
}
// Per-level metric ratios of the normalised gradients (velocity_tracer_gradients.jl:126-260: Δᶠxᶜᶜᶜ = 2Δx etc.), tabulated on the
// host with the reference's own IEEE divisions: the kernel would otherwise evaluate 18 Float64 divisions per cell (a third of
// its instructions) to re-derive numbers that depend on the level only.  Row r of `tab` holds Nz + 2 values for k = 0 .. Nz+1:
//   0: fz = 2 Δzᶜ(k)   1: fx / fz   2: fz / fx   3: fy / fz   4: fz / fy   5: Δ² = 3 / (1/fx² + 1/fy² + 1/fz²)
template <typename T>
struct AmdGeom {
    const T *tab;
    int stride;      // Nz + 2
    T rxy, ryx;      // fx / fy, fy / fx
    __device__ __forceinline__ T at(int r, int k) const { return __ldg(tab + r * stride + k); }
};
// MINB: resident CTAs per SM the register budget is set for (4: 126 registers, no spills; 5: 96 registers, 0.3 KB of spills;
// the host picks it -- OB_AMD_MINB in the environment overrides, for A/B timings)
template <typename T, int MINB>
__global__ void __launch_bounds__(128, MINB) amd_kernel(const __grid_constant__ TendP<T> P, int m, const AmdGeom<T> A) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    const GridD<T> &g = P.g;
    const ClosureD<T> &cl = P.cl[m];
    const int ox = g.topo[0] == FLAT ? 0 : 1;
    // fields may differ in their strides (Face fields on Bounded directions): one base pointer per field, then 32-bit
    // relative offsets (a, b, c are compile-time after unrolling)
    struct Cur { const T *p; int oy, oz; };
    auto cur = [&](const Fld<T> &f) {
        Cur c;
        c.p = f.p + (f.off + i + (long)j * f.sy + (long)k * f.sz);
        c.oy = g.topo[1] == FLAT ? 0 : f.sy;
        c.oz = g.topo[2] == FLAT ? 0 : (int)f.sz;
        return c;
    };
    const Cur cu = cur(P.u), cv = cur(P.v), cw = cur(P.w);
    auto at = [&](const Cur &f, int a, int b, int c) -> T { return __ldg(f.p + (a * ox + b * f.oy + c * f.oz)); };
    const T rdx = g.rdx, rdy = g.rdy;
    const T fx = 2 * g.dx, fy = 2 * g.dy;
    T fz[2], rdzf[2], rxz[2], rzx[2], ryz[2], rzy[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int kc = g.topo[2] == FLAT ? k : k + c;
        rdzf[c] = g.rdzF(kc);
        fz[c] = A.at(0, kc); rxz[c] = A.at(1, kc); rzx[c] = A.at(2, kc); ryz[c] = A.at(3, kc); rzy[c] = A.at(4, kc);
    }
    const T rxy = A.rxy, ryx = A.ryx;
    const T delta2 = A.at(5, k);
    // ccc
    const T ux = (at(cu, 1, 0, 0) - at(cu, 0, 0, 0)) * rdx;
    const T vy = (at(cv, 0, 1, 0) - at(cv, 0, 0, 0)) * rdy;
    const T wz = (at(cw, 0, 0, 1) - at(cw, 0, 0, 0)) * g.rdzC(k);
    // ffc [a][b], fcf [a][c], cff [b][c]
    T xv[2][2], yu[2][2], xw[2][2], zu[2][2], yw[2][2], zv[2][2], s12[2][2], s13[2][2], s23[2][2];
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
        for (int p = 0; p < 2; p++) {
            xv[p][q] = rxy * ((at(cv, p, q, 0) - at(cv, p - 1, q, 0)) * rdx);
            yu[p][q] = ryx * ((at(cu, p, q, 0) - at(cu, p, q - 1, 0)) * rdy);
            s12[p][q] = T(0.5) * (yu[p][q] + xv[p][q]);
            xw[p][q] = rxz[q] * ((at(cw, p, 0, q) - at(cw, p - 1, 0, q)) * rdx);
            zu[p][q] = rzx[q] * ((at(cu, p, 0, q) - at(cu, p, 0, q - 1)) * rdzf[q]);
            s13[p][q] = T(0.5) * (zu[p][q] + xw[p][q]);
            yw[p][q] = ryz[q] * ((at(cw, 0, p, q) - at(cw, 0, p - 1, q)) * rdy);
            zv[p][q] = rzy[q] * ((at(cv, 0, p, q) - at(cv, 0, p, q - 1)) * rdzf[q]);
            s23[p][q] = T(0.5) * (zv[p][q] + yw[p][q]);
        }
    auto I = [&](auto fn) {  // interp4 of a pointwise expression fn(p, q)
        T t[2][2];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int p = 0; p < 2; p++) t[p][q] = fn(p, q);
        return interp4<T>(t);
    };
#define SQ(A) I([&](int p, int q) { return A[p][q] * A[p][q]; })
#define PR(A, B) I([&](int p, int q) { return A[p][q] * B[p][q]; })
#define ID(A) I([&](int p, int q) { return A[p][q]; })
    const T I_xv2 = SQ(xv), I_yu2 = SQ(yu), I_xw2 = SQ(xw), I_zu2 = SQ(zu), I_yw2 = SQ(yw), I_zv2 = SQ(zv);
    const T I_xv = ID(xv), I_yu = ID(yu), I_xw = ID(xw), I_zu = ID(zu), I_yw = ID(yw), I_zv = ID(zv);
    {
        const T q = ux * ux + vy * vy + wz * wz + I_xv2 + I_yu2 + I_xw2 + I_zu2 + I_yw2 + I_zv2;
        T nu = 0;
        if (q != 0) {
            const T r1 = ux * (ux * ux) + vy * I_xv2 + wz * I_xw2 + 2 * ux * PR(xv, s12) + 2 * ux * PR(xw, s13) + 2 * I_xv * I_xw * ID(s23);
            const T r2 = ux * I_yu2 + vy * (vy * vy) + wz * I_yw2 + 2 * vy * PR(yu, s12) + 2 * I_yu * I_yw * ID(s13) + 2 * vy * PR(yw, s23);
            const T r3 = ux * I_zu2 + vy * I_zv2 + wz * (wz * wz) + 2 * I_zu * I_zv * ID(s12) + 2 * wz * PR(zu, s13) + 2 * wz * PR(zv, s23);
            const T r = r1 + r2 + r3;
            T cbz = 0;
            if (cl.amd_has_cb) {  // Cb_norm_wᵢ_bᵢᶜᶜᶜ (:320-333): ℑ of ∂b at the faces (bpert needs absolute indices)
                auto bp = [&](int a, int b, int c) { return bpert(P, i + a * ox, j + (g.topo[1] == FLAT ? 0 : b), k + (g.topo[2] == FLAT ? 0 : c)); };
                const T bx = T(0.5) * ((bp(0, 0, 0) - bp(-1, 0, 0)) * rdx + (bp(1, 0, 0) - bp(0, 0, 0)) * rdx);
                const T by = T(0.5) * ((bp(0, 0, 0) - bp(0, -1, 0)) * rdy + (bp(0, 1, 0) - bp(0, 0, 0)) * rdy);
                const T bz = T(0.5) * ((bp(0, 0, 0) - bp(0, 0, -1)) * rdzf[0] + (bp(0, 0, 1) - bp(0, 0, 0)) * rdzf[1]);
                cbz = cl.cb * (I_xw * fx * bx + I_yw * fy * by + wz * fz[0] * bz);
            }
            cbz = cbz / fz[0];
            nu = -cl.Cnu * delta2 * (r - cbz) / q;
        }
        P.nue[m](i, j, k) = fmax(T(0), nu);
    }
    for (int t = 0; t < P.ntr; t++) {
        const Cur c = cur(P.c[t]);
        T xc[2], yc[2], zc[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            xc[p] = fx * ((at(c, p, 0, 0) - at(c, p - 1, 0, 0)) * rdx);
            yc[p] = fy * ((at(c, 0, p, 0) - at(c, 0, p - 1, 0)) * rdy);
            zc[p] = fz[p] * ((at(c, 0, 0, p) - at(c, 0, 0, p - 1)) * rdzf[p]);
        }
        const T icx2 = T(0.5) * (xc[0] * xc[0] + xc[1] * xc[1]), icy2 = T(0.5) * (yc[0] * yc[0] + yc[1] * yc[1]), icz2 = T(0.5) * (zc[0] * zc[0] + zc[1] * zc[1]);
        const T sigma = icx2 + icy2 + icz2;
        T kap_ = 0;
        if (sigma != 0) {
            const T icx = T(0.5) * (xc[0] + xc[1]), icy = T(0.5) * (yc[0] + yc[1]), icz = T(0.5) * (zc[0] + zc[1]);
            // cy_uy uses ℑxzᶜᵃᶜ of norm_∂y_w exactly as the reference does (:345): norm_∂y_w at (i+a, j, k+c)
            T yw_xz[2][2];
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int p = 0; p < 2; p++) yw_xz[p][q] = ryz[q] * ((at(cw, p, 0, q) - at(cw, p, -1, q)) * rdy);
            const T cx = ux * icx2 + I_xv * icx * icy + I_xw * icx * icz;
            const T cy = I_yu * icy * icx + vy * icy2 + interp4<T>(yw_xz) * icy * icz;
            const T cz = I_zu * icz * icx + I_zv * icz * icy + wz * icz2;
            const T theta = cx + cy + cz;
            kap_ = -cl.Ckappa[t] * delta2 * theta / sigma;
        }
        P.kappae[m][t](i, j, k) = fmax(T(0), kap_);
    }
#undef SQ
#undef PR
#undef ID
}

}  // namespace ob
