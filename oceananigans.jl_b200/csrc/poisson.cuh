// poisson.cuh -- pressure solvers: FFTBasedPoissonSolver (src/Solvers/fft_based_poisson_solver.jl:51-124) and
// FourierTridiagonalPoissonSolver with the BatchedTridiagonalSolver Thomas sweep
// (fourier_tridiagonal_poisson_solver.jl:87-260, batched_tridiagonal_solver.jl:211-243).
//
// cuFFT does only the batched 1-D / 2-D complex transforms.  Hand-written kernels do everything around them:
// the Makhoul (1980) DCT index permutation fused with the load, the twiddle multiply fused with the store
// (reference: 5-6 separate full-array passes per Bounded dimension, discrete_transforms.jl:114-183), the
// eigenvalue division, and the complex Thomas sweep fused with the mean removal.
#pragma once
#include <cufft.h>
#include <math.h>
#include <vector>
#include "common.cuh"

namespace ob {

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float> { using type = float2; };

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Makhoul permutation (index_permutations.jl:18-36), 0-based: even i -> i/2 ; odd i -> N-1-(i-1)/2
__device__ __forceinline__ int makhoul(int i, int N) { return (i & 1) ? N - 1 - (i - 1) / 2 : i / 2; }

// Layout of the work array B of the y transform: `ty` = 1 stores it with y fastest, [k][i][j], so that the 1-D FFT along y is ONE
// contiguous batched cuFFT call over all Nx Nz lines (in the natural layout a y line has stride Nx and cuFFT's single batch
// dimension covers one z plane per call: Nz launches per transform).  The permute / twiddle passes around the FFT exist anyway;
// they read or write B through this index.
__device__ __forceinline__ long bidx(int i, int j, int k, int Nx, int Ny, int ty) {
    return ty ? j + (long)Ny * (i + (long)Nx * k) : i + (long)Nx * (j + (long)Ny * k);
}
// plain change of layout (Periodic y outside the 2-D (x, y) plan): B[k][i][j] = A[k][j][i] and back
template <typename C>
__global__ void __launch_bounds__(256) transpose_y_kernel(const C *__restrict__ A, C *__restrict__ B, int Nx, int Ny, int Nz, int back) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    if (back) B[t] = A[bidx(i, j, k, Nx, Ny, 1)];
    else B[bidx(i, j, k, Nx, Ny, 1)] = A[t];
}
// B[perm_d(idx)] = A[idx] along dimension d
template <typename C>
__global__ void __launch_bounds__(256) permute_kernel(const C *__restrict__ A, C *__restrict__ B, int Nx, int Ny, int Nz, int d, int ty = 0) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    int ii = i, jj = j, kk = k;
    if (d == 0) ii = makhoul(i, Nx); else if (d == 1) jj = makhoul(j, Ny); else kk = makhoul(k, Nz);
    B[bidx(ii, jj, kk, Nx, Ny, ty)] = A[t];
}
// forward twiddle: A = 2 * real(ω_4N^k * B)   (discrete_transforms.jl:172-175; twiddles :48-78)
template <typename T, typename C>
__global__ void __launch_bounds__(256) twiddle_fwd_kernel(const C *__restrict__ B, C *__restrict__ A, const C *__restrict__ w,
                                                          int Nx, int Ny, int Nz, int d, int ty = 0) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    int q = d == 0 ? i : d == 1 ? j : k;
    C v = cmul(w[q], B[ty ? bidx(i, j, k, Nx, Ny, 1) : t]);
    C o;
    o.x = 2 * v.x;
    o.y = 0;
    A[t] = o;
}
// backward twiddle: B = A * ω_4N^{-k} (k = 0 halved)   (discrete_transforms.jl:177-183)
template <typename T, typename C>
__global__ void __launch_bounds__(256) twiddle_bwd_kernel(const C *__restrict__ A, C *__restrict__ B, const C *__restrict__ w,
                                                          int Nx, int Ny, int Nz, int d, int ty = 0) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    int q = d == 0 ? i : d == 1 ? j : k;
    B[ty ? bidx(i, j, k, Nx, Ny, 1) : t] = cmul(A[t], w[q]);
}
// unpermute + real: A[idx] = real(B[perm_d(idx)])
template <typename C>
__global__ void __launch_bounds__(256) unpermute_kernel(const C *__restrict__ B, C *__restrict__ A, int Nx, int Ny, int Nz, int d, int ty = 0) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    int ii = i, jj = j, kk = k;
    if (d == 0) ii = makhoul(i, Nx); else if (d == 1) jj = makhoul(j, Ny); else kk = makhoul(k, Nz);
    C v = B[bidx(ii, jj, kk, Nx, Ny, ty)];
    v.y = 0;
    A[t] = v;
}
// ϕ̂ = -b̂ / (λx + λy + λz) ; ϕ̂[1,1,1] = 0   (fft_based_poisson_solver.jl:109-114)
template <typename T, typename C>
__global__ void __launch_bounds__(256) eigen_divide_kernel(C *__restrict__ A, const T *__restrict__ lx, const T *__restrict__ ly,
                                                           const T *__restrict__ lz, int Nx, int Ny, int Nz) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % Nx), j = (int)((t / Nx) % Ny), k = (int)(t / ((long)Nx * Ny));
    C v = A[t];
    T lam = lx[i] + ly[j] + lz[k];
    C o;
    if (t == 0) { o.x = 0; o.y = 0; }
    else { o.x = -v.x / lam; o.y = -v.y / lam; }
    A[t] = o;
}

// Real-data DCT path (x, y Periodic, z Bounded and regular): the rhs is real, so the z transform is a real-to-complex FFT
// of the Makhoul-permuted line and the DCT-II coefficients X[k] = 2 Re(ω_k V[k]) are REAL; V[k] for k > N/2 comes from
// the conjugate symmetry V[N-k] = conj(V[k]).  The (x, y) transform is then a real-to-complex 2-D FFT of X.  Backward:
// V[k] = ω⁻_k (X[k] - i X[N-k]) / 2 (X[N] = 0) is the Hermitian half spectrum whose complex-to-real FFT is the permuted
// line.  Same transforms as REDFT10 / REDFT01 (plan_transforms.jl:16-34), a quarter of the bytes of the complex route.
template <typename T, typename C>
__global__ void __launch_bounds__(256) dct_z_real_fwd_kernel(const C *__restrict__ V, T *__restrict__ X, const C *__restrict__ w,
                                                             int Nx, int Ny, int Nz, int Nzh) {
    const long n = (long)Nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long plane = (long)Nx * Ny;
    const int k = (int)(t / plane);
    const long ij = t - (long)k * plane;
    (void)Nzh;
    C v;
    if (2 * k <= Nz) v = V[ij + (long)k * plane];
    else { v = V[ij + (long)(Nz - k) * plane]; v.y = -v.y; }
    const C r = cmul(w[k], v);
    X[t] = 2 * r.x;
}
template <typename T, typename C>
__global__ void __launch_bounds__(256) dct_z_real_bwd_kernel(const T *__restrict__ X, C *__restrict__ V, const C *__restrict__ w,
                                                             int Nx, int Ny, int Nz, int Nzh) {
    const long n = (long)Nx * Ny * Nzh;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long plane = (long)Nx * Ny;
    const int k = (int)(t / plane);
    const long ij = t - (long)k * plane;
    const T a = X[ij + (long)k * plane];
    const T b = k == 0 ? T(0) : X[ij + (long)(Nz - k) * plane];
    C z; z.x = a; z.y = -b;                 // X[k] - i X[N-k]
    V[t] = cmul(w[k], z);                    // w[k] = ω⁻_k / 2
}

// Complex Thomas sweep along z, one thread per (i,j) column, coalesced along x
// (batched_tridiagonal_solver.jl:217-243).  a = c = lower diagonal (Nz-1), D = main diagonal (Nx,Ny,Nz),
// tscr = real scratch (Nx,Ny,Nz).  In place on the complex array A (f and ϕ alias).
// Column (1,1) additionally removes its k-mean, which equals `ϕ .-= mean(ϕ)` of
// fourier_tridiagonal_poisson_solver.jl:252-254 (only the horizontal-mean mode has a non-zero mean).
// un-contracted arithmetic (the reference CPU code has no @muladd in the Thomas sweep; β of the singular
// column is a pure rounding residue, so contraction would change it at O(1))
template <typename T, typename C>
__global__ void __launch_bounds__(128) thomas_kernel(C *__restrict__ A, const T *__restrict__ lower, const T *__restrict__ upper,
                                                     const T *__restrict__ D, T *__restrict__ tscr, int Nx, int Ny, int Nz, T eps10,
                                                     int remove_mean) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= Nx) return;
    const long s = (long)Nx * Ny;
    const long o = i + (long)Nx * j;
    T beta = D[o];
    C f = A[o];
    C phi;
    phi.x = f.x / beta;
    phi.y = f.y / beta;
    A[o] = phi;
    // The recurrences are serial in k but their operands are not: the diagonal and the right-hand side of the NEXT four levels are
    // loaded while the current four are being eliminated (the sweep is latency-bound otherwise: one dependent miss per level).
    constexpr int PF = 4;
    T dn[PF];
    C fn[PF];
#pragma unroll
    for (int q = 0; q < PF; q++) {
        const int k = 1 + q;
        if (k < Nz) { dn[q] = D[o + k * s]; fn[q] = A[o + k * s]; }
    }
    for (int k0 = 1; k0 < Nz; k0 += PF) {
        T dc[PF];
        C fc[PF];
#pragma unroll
        for (int q = 0; q < PF; q++) { dc[q] = dn[q]; fc[q] = fn[q]; }
#pragma unroll
        for (int q = 0; q < PF; q++) {
            const int k = k0 + PF + q;
            if (k < Nz) { dn[q] = D[o + k * s]; fn[q] = A[o + k * s]; }
        }
#pragma unroll
        for (int q = 0; q < PF; q++) {
            const int k = k0 + q;
            if (k >= Nz) break;
            const T cm = upper[k - 1], am = lower[k - 1];
            const T tk = cm / beta;
            tscr[o + k * s] = tk;
            beta = sub_rn(dc[q], mul_rn(am, tk));
            f = fc[q];
            const bool ok = fabs(beta) > eps10;
            C star;
            star.x = sub_rn(f.x, mul_rn(am, phi.x)) / beta;
            star.y = sub_rn(f.y, mul_rn(am, phi.y)) / beta;
            // not definitely diagonally dominant (the singular λ = 0 column): the reference leaves ϕ[k] untouched, i.e.
            // an arbitrary stale value that only shifts the null-space constant removed by the mean subtraction; 0 here.
            if (!ok) { star.x = 0; star.y = 0; }
            phi = star;
            A[o + k * s] = phi;
        }
    }
    // back substitution, operands of the next four levels in flight likewise
    T tn[PF];
    C an[PF];
#pragma unroll
    for (int q = 0; q < PF; q++) {
        const int k = Nz - 2 - q;
        if (k >= 0) { tn[q] = tscr[o + (k + 1) * s]; an[q] = A[o + k * s]; }
    }
    for (int k0 = Nz - 2; k0 >= 0; k0 -= PF) {
        T tc[PF];
        C ac[PF];
#pragma unroll
        for (int q = 0; q < PF; q++) { tc[q] = tn[q]; ac[q] = an[q]; }
#pragma unroll
        for (int q = 0; q < PF; q++) {
            const int k = k0 - PF - q;
            if (k >= 0) { tn[q] = tscr[o + (k + 1) * s]; an[q] = A[o + k * s]; }
        }
#pragma unroll
        for (int q = 0; q < PF; q++) {
            const int k = k0 - q;
            if (k < 0) break;
            C cur = ac[q];
            cur.x = sub_rn(cur.x, mul_rn(tc[q], phi.x));
            cur.y = sub_rn(cur.y, mul_rn(tc[q], phi.y));
            phi = cur;
            A[o + k * s] = phi;
        }
    }
    if (remove_mean && o == 0) {
        T mx = 0, my = 0;
        for (int k = 0; k < Nz; k++) { C v = A[k * s]; mx += v.x; my += v.y; }
        mx /= Nz; my /= Nz;
        for (int k = 0; k < Nz; k++) { C v = A[k * s]; v.x -= mx; v.y -= my; A[k * s] = v; }
    }
}

}  // namespace ob
