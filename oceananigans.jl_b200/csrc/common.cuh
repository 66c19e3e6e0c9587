// common.cuh -- device-side descriptors shared by all kernels of libocean_b200.so.
//
// Data layout in HBM is the reference's: each field is ONE contiguous parent array, x fastest,
// of size (Nx+2Hx, Ny+2Hy, Nz+2Hz) (+1 along a Bounded direction for Face-located fields; size N with no
// halo along Flat directions) -- src/Grids/new_data.jl:11-74.  Logical (1-based, Julia) indices are used
// throughout so that every expression can be checked against the reference line it restates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define OB_MAXTR 8
#define OB_MAXCL 4
#define OB_MAXBUF 6

namespace ob {

enum { PERIODIC = 0, BOUNDED = 1, FLAT = 2 };
enum { ADV_NONE = 0, ADV_CENTERED = 1, ADV_WENO = 2 };
enum { CL_SCALAR = 1, CL_SMAG = 2, CL_AMD = 3 };
enum { BUOY_NONE = 0, BUOY_TRACER = 1, BUOY_SEAWATER = 2 };
enum { BC_NONE = 0, BC_PERIODIC = 1, BC_FLUX = 2, BC_VALUE = 3, BC_GRADIENT = 4, BC_IMPENETRABLE = 5, BC_COMM = 6 };

// Field view: value at logical (i,j,k) = p[off + i + j*sy + k*sz]
template <typename T>
struct Fld {
    T *p;
    long off;   // (ox-1) + (oy-1)*sy + (oz-1)*sz  with o* = parent index of logical index 1
    int sy;     // Px
    long sz;    // Px*Py
    __device__ __forceinline__ T &operator()(int i, int j, int k) const { return p[off + i + (long)j * sy + (long)k * sz]; }
    __device__ __forceinline__ T ld(int i, int j, int k) const { return __ldg(p + (off + i + (long)j * sy + (long)k * sz)); }
    __device__ __forceinline__ long idx(int i, int j, int k) const { return off + i + (long)j * sy + (long)k * sz; }
};

// RectilinearGrid with regular x, y and regular-or-stretched z (the FFT / Fourier-tridiagonal pressure solvers
// require exactly this: src/Models/NonhydrostaticModels/NonhydrostaticModels.jl:30-52).
template <typename T>
struct GridD {
    int N[3], H[3], topo[3];
    T dx, dy, dz;        // regular spacings (1 for Flat: spacings_and_areas_and_volumes.jl:138)
    const T *dzf;        // stretched z: Δzᵃᵃᶠ, pre-offset so dzf[k] is logical k; nullptr when regular
    const T *dzc;        // stretched z: Δzᵃᵃᶜ
    // reciprocals, precomputed on the host with the same IEEE division the reference evaluates per call
    // (Δ⁻¹ = 1/Δ, reciprocal_metric_operators.jl:13-26; V⁻¹ = 1/((Δx Δy) Δz), spacings_and_areas_and_volumes.jl:483-491)
    T rdx, rdy, rdz;     // 1/Δx, 1/Δy, 1/Δz (regular)
    T rvol;              // 1/((Δx Δy) Δz) (regular z)
    const T *rdzf, *rdzc;  // stretched z: 1/Δzᶠ(k), 1/Δzᶜ(k) (pre-offset, logical k)
    const T *rvf, *rvc;    // stretched z: 1/((Δx Δy) Δzᶠ(k)), 1/((Δx Δy) Δzᶜ(k))
    __device__ __forceinline__ T dzF(int k) const { return dzf ? __ldg(dzf + k) : dz; }
    __device__ __forceinline__ T dzC(int k) const { return dzc ? __ldg(dzc + k) : dz; }
    __device__ __forceinline__ T rdzF(int k) const { return dzf ? __ldg(rdzf + k) : rdz; }
    __device__ __forceinline__ T rdzC(int k) const { return dzc ? __ldg(rdzc + k) : rdz; }
    __device__ __forceinline__ T rVf(int k) const { return dzf ? __ldg(rvf + k) : rvol; }
    __device__ __forceinline__ T rVc(int k) const { return dzc ? __ldg(rvc + k) : rvol; }
};

template <typename T>
struct ClosureD {
    int kind;
    T nu;
    T kappa[OB_MAXTR];
    T Pr[OB_MAXTR];
    T cs, cb;
    int lilly;
    T Cnu;
    T Ckappa[OB_MAXTR];
    int amd_has_cb;
    int vi;   // VerticallyImplicitTimeDiscretization: interior vertical fluxes are elided from the explicit tendencies
};

template <typename T>
struct BcD {
    int kind[6];
    T value[6];
};

// un-contracted arithmetic for the places where the reference CPU code has no @muladd and the result must be
// bit-identical (halo extrapolation) or is a pure rounding residue (Thomas sweep of the singular column)
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// ℑxy of four area-weighted values, 0.5 (0.5 (A a + A b) + 0.5 (A c + A d)), evaluated as the reference does: products and sums
// rounded one by one (interpolation_operators.jl has no @muladd).  `A a + A b` would otherwise contract into an FMA around
// either product, at the compiler's whim, and two kernels of this library would differ in the last bit.
template <typename T> __device__ __forceinline__ T interp4_rn(T A, T a, T b, T c, T d) {
    const T h = T(0.5);
    return mul_rn(h, add_rn(mul_rn(h, add_rn(mul_rn(A, a), mul_rn(A, b))), mul_rn(h, add_rn(mul_rn(A, c), mul_rn(A, d)))));
}
template <typename T> __device__ __forceinline__ T interp2_rn(T A, T a, T b) { return mul_rn(T(0.5), add_rn(mul_rn(A, a), mul_rn(A, b))); }

// Makhoul (1980) permutation of a DCT line (index_permutations.jl:18-36), 0-based: even i -> i/2 ; odd i -> N-1-(i-1)/2
__device__ __forceinline__ int makhoul_index(int i, int N) { const int h = i >> 1; return (i & 1) ? N - 1 - h : h; }  // i >= 0

// Flattened launch geometry: blockIdx.x enumerates (x-block, j, k); returns false for the x tail.
__device__ __forceinline__ bool cell_from_block(int Nx, int Ny, int &i, int &j, int &k) {
    const int nbx = (Nx + blockDim.x - 1) / blockDim.x;
    const long b = blockIdx.x;
    const int bx = (int)(b % nbx);
    const long jk = b / nbx;
    i = 1 + bx * blockDim.x + threadIdx.x;
    j = 1 + (int)(jk % Ny);
    k = 1 + (int)(jk / Ny);
    return i <= Nx;
}

}  // namespace ob
