// dist_solver.cuh -- slab-x distributed pressure solvers (one process per GPU, NCCL over NVLink/NVSwitch).
//
// Reference: src/DistributedComputations/distributed_fft_based_poisson_solver.jl:141-183 (z-FFT, [z->y no-op for
// slabs], y-FFT, y->x all-to-all, x-FFT, eigenvalue division, and back), distributed_fft_tridiagonal_solver.jl:292-316
// (stretched z: the Thomas sweep runs in the x-local layout), transposable_field.jl:49-110 and
// distributed_transpose.jl:39-109 (pack / unpack index maps), ext/OceananigansNCCLExt/nccl_transpose.jl:47-75 (grouped
// ncclSend/ncclRecv all-to-all).  Included by ocean_b200.cu after SolverT.
//
// Layouts (x fastest everywhere): S  = y-local "slab" layout   (nx, Ny, Nz),  nx = Nx/R
//                                 Tt = x-local transposed layout (Nx, ny, Nz), ny = Ny/R
// Peer chunk r of a transpose buffer holds (nx, ny, Nz) elements: [xl + nx*(yl + ny*z)].
#pragma once
#include <nccl.h>

#define NCCL_TRY(x)                                                                                       \
    do {                                                                                                  \
        ncclResult_t r_ = (x);                                                                            \
        if (r_ != ncclSuccess) return fail(OB_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #x, ncclGetErrorString(r_)); \
    } while (0)

namespace ob {

// slab -> per-peer chunks (integer index work: bit-exact)
template <typename C>
__global__ void __launch_bounds__(256) pack_y_to_x_kernel(const C *__restrict__ S, C *__restrict__ buf, int nx, int ny, int Ny, int Nz) {
    const long n = (long)nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int xl = (int)(t % nx), y = (int)((t / nx) % Ny), z = (int)(t / ((long)nx * Ny));
    const int r = y / ny, yl = y - r * ny;
    buf[(long)r * nx * ny * Nz + xl + (long)nx * (yl + (long)ny * z)] = S[t];
}
// received chunks -> transposed layout
template <typename C>
__global__ void __launch_bounds__(256) unpack_y_to_x_kernel(const C *__restrict__ buf, C *__restrict__ Tt, int nx, int ny, int NxG, int Nz) {
    const long n = (long)NxG * ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = (int)(t % NxG), yl = (int)((t / NxG) % ny), z = (int)(t / ((long)NxG * ny));
    const int r = x / nx, xl = x - r * nx;
    Tt[t] = buf[(long)r * nx * ny * Nz + xl + (long)nx * (yl + (long)ny * z)];
}
template <typename C>
__global__ void __launch_bounds__(256) pack_x_to_y_kernel(const C *__restrict__ Tt, C *__restrict__ buf, int nx, int ny, int NxG, int Nz) {
    const long n = (long)NxG * ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = (int)(t % NxG), yl = (int)((t / NxG) % ny), z = (int)(t / ((long)NxG * ny));
    const int r = x / nx, xl = x - r * nx;
    buf[(long)r * nx * ny * Nz + xl + (long)nx * (yl + (long)ny * z)] = Tt[t];
}
template <typename C>
__global__ void __launch_bounds__(256) unpack_x_to_y_kernel(const C *__restrict__ buf, C *__restrict__ S, int nx, int ny, int Ny, int Nz) {
    const long n = (long)nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int xl = (int)(t % nx), y = (int)((t / nx) % Ny), z = (int)(t / ((long)nx * Ny));
    const int r = y / ny, yl = y - r * ny;
    S[t] = buf[(long)r * nx * ny * Nz + xl + (long)nx * (yl + (long)ny * z)];
}
// x <-> z transposition (used when Nz % R == 0 and the solver is FFT-based): the slab layout (nx, Ny, Nz) is z-slowest,
// so the chunk for peer r -- levels r*nz .. (r+1)*nz-1 -- is CONTIGUOUS and is sent straight from S without packing;
// the receiver scatters rows of nx into the z-local layout Tz = (Nx, Ny, nz), where the (x, y) transform is one
// contiguous batched 2-D cuFFT exactly as on a single GPU.
template <typename C>
__global__ void __launch_bounds__(256) unpack_z_to_x_kernel(const C *__restrict__ buf, C *__restrict__ Tz, int nx, int NxG, int Ny, int nz) {
    const long n = (long)NxG * Ny * nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = (int)(t % NxG), y = (int)((t / NxG) % Ny), zl = (int)(t / ((long)NxG * Ny));
    const int r = x / nx, xl = x - r * nx;
    Tz[t] = buf[(long)r * nx * Ny * nz + xl + (long)nx * (y + (long)Ny * zl)];
}
template <typename C>
__global__ void __launch_bounds__(256) pack_x_to_z_kernel(const C *__restrict__ Tz, C *__restrict__ buf, int nx, int NxG, int Ny, int nz) {
    const long n = (long)NxG * Ny * nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = (int)(t % NxG), y = (int)((t / NxG) % Ny), zl = (int)(t / ((long)NxG * Ny));
    const int r = x / nx, xl = x - r * nx;
    buf[(long)r * nx * Ny * nz + xl + (long)nx * (y + (long)Ny * zl)] = Tz[t];
}
// Fused transposition over NVLink peer memory: every rank PULLS its part of the transposed array straight out of the
// peers' storage (CUDA IPC mappings; rows of nx contiguous complex numbers), so the all-to-all and the unpack are ONE
// kernel and no staging buffer is touched.  Ordering across GPUs is provided by a stream-ordered barrier before each pull.
#define OB_MAX_PEERS 16
template <typename C> struct PeerPtrs { const C *p[OB_MAX_PEERS]; };
template <typename C>
__global__ void __launch_bounds__(256) pull_z_to_x_kernel(const __grid_constant__ PeerPtrs<C> peers, C *__restrict__ Tz, int nx, int NxG, int Ny,
                                                          int nz, int Nz, int rank) {
    const long n = (long)NxG * Ny * nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int x = (int)(t % NxG), y = (int)((t / NxG) % Ny), zl = (int)(t / ((long)NxG * Ny));
    const int r = x / nx, xl = x - r * nx;
    (void)Nz;
    Tz[t] = peers.p[r][xl + (long)nx * (y + (long)Ny * (rank * nz + zl))];   // peer r's slab layout (nx, Ny, Nz)
}
template <typename C>
__global__ void __launch_bounds__(256) pull_x_to_z_kernel(const __grid_constant__ PeerPtrs<C> peers, C *__restrict__ S, int nx, int NxG, int Ny,
                                                          int nz, int Nz, int rank) {
    const long n = (long)nx * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    // the slowest index walks the peers starting from this rank, so that at any moment the ranks read from DIFFERENT peers
    // (walking 0 .. R-1 everywhere makes all of them pull on one peer's links at a time)
    const int xl = (int)(t % nx), y = (int)((t / nx) % Ny), q = (int)(t / ((long)nx * Ny));
    const int R = Nz / nz;
    int r = rank + q / nz;
    if (r >= R) r -= R;
    const int zl = q % nz;
    S[xl + (long)nx * (y + (long)Ny * (r * nz + zl))] = peers.p[r][(rank * nx + xl) + (long)NxG * (y + (long)Ny * zl)];     // peer r's z-local layout (Nx, Ny, nz)
}

// the same with one extra column in front of every row: column 0 of the output is global column rank*nx - 1 (periodic), the
// last column of the west neighbour's slab -- every peer holds complete x lines in the z-local layout, so the projection's
// west neighbour of p comes along with the transposition instead of through a halo exchange
template <typename C>
__global__ void __launch_bounds__(256) pull_x_to_z_west_kernel(const __grid_constant__ PeerPtrs<C> peers, C *__restrict__ S, int nx, int NxG, int Ny,
                                                               int nz, int Nz, int rank) {
    const int nx1 = nx + 1;
    const long n = (long)nx1 * Ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int xl = (int)(t % nx1), y = (int)((t / nx1) % Ny), q = (int)(t / ((long)nx1 * Ny));
    const int R = Nz / nz;
    int r = rank + q / nz;   // peers in rotated order (see pull_x_to_z_kernel)
    if (r >= R) r -= R;
    const int zl = q % nz;
    int xg = rank * nx + xl - 1;
    if (xg < 0) xg += NxG;
    S[xl + (long)nx1 * (y + (long)Ny * (r * nz + zl))] = peers.p[r][xg + (long)NxG * (y + (long)Ny * zl)];
}

// Level-chunked forms of the two pulls and of the eigenvalue division (pipelined solve: the transposition of one chunk of
// local levels overlaps the (x, y) transforms of another).  Levels zl0 .. zl0 + len - 1 of the z-local layout.
// Both are grid-stride kernels launched with a bounded number of CTAs (4 loads in flight per thread): they run on the low-
// priority stream next to the transforms of the main stream and must not fill every CTA slot of the device.
template <typename C>
__global__ void __launch_bounds__(256) pull_z_to_x_chunk_kernel(const __grid_constant__ PeerPtrs<C> peers, C *__restrict__ Tz, int nx, int NxG, int Ny,
                                                                int nz, int rank, int zl0, int len) {
    const long plane = (long)NxG * Ny, n = plane * len, stride = (long)gridDim.x * blockDim.x;
    for (long t0 = (long)blockIdx.x * blockDim.x + threadIdx.x; t0 < n; t0 += 4 * stride) {
        C v[4];
        long dst[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const long t = t0 + q * stride;
            dst[q] = -1;
            if (t < n) {
                const int x = (int)(t % NxG), y = (int)((t / NxG) % Ny), zl = zl0 + (int)(t / plane);
                const int r = x / nx, xl = x - r * nx;
                dst[q] = x + (long)NxG * (y + (long)Ny * zl);
                v[q] = peers.p[r][xl + (long)nx * (y + (long)Ny * (rank * nz + zl))];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (dst[q] >= 0) Tz[dst[q]] = v[q];
    }
}
// rows of nx + west columns (west = 1: the west neighbour's last column in front, see pull_x_to_z_west_kernel)
template <typename C>
__global__ void __launch_bounds__(256) pull_x_to_z_chunk_kernel(const __grid_constant__ PeerPtrs<C> peers, C *__restrict__ S, int nx, int NxG, int Ny,
                                                                int nz, int R, int rank, int west, int zl0, int len) {
    const int nxo = nx + west;
    const long row = (long)nxo * Ny, n = row * len * R, stride = (long)gridDim.x * blockDim.x;
    for (long t0 = (long)blockIdx.x * blockDim.x + threadIdx.x; t0 < n; t0 += 4 * stride) {
        C v[4];
        long dst[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            const long t = t0 + q4 * stride;
            dst[q4] = -1;
            if (t < n) {
                const int xl = (int)(t % nxo), y = (int)((t / nxo) % Ny);
                const int q = (int)(t / row);            // (peer in rotated order, level of the chunk)
                int r = rank + q / len;
                if (r >= R) r -= R;
                const int zl = zl0 + q % len;
                int xg = rank * nx + xl - west;
                if (xg < 0) xg += NxG;
                dst[q4] = xl + (long)nxo * (y + (long)Ny * (r * nz + zl));
                v[q4] = peers.p[r][xg + (long)NxG * (y + (long)Ny * zl)];
            }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) if (dst[q4] >= 0) S[dst[q4]] = v[q4];
    }
}

// Stream-ordered barrier across the ranks without a collective: lane r publishes this rank's epoch into peer r's flag
// array (release store over NVLink) and then waits until peer r's epoch has arrived in the local array.
struct PeerFlags { int *p[OB_MAX_PEERS]; };
__global__ void ipc_barrier_kernel(const __grid_constant__ PeerFlags peers, int *mine, int R, int rank, int epoch) {
    const int r = threadIdx.x;
    if (r < R) {
        __threadfence_system();
        st_release_sys(peers.p[r] + rank, epoch);
        while (ld_acquire_sys(mine + r) < epoch) {}
    }
}

template <typename T, typename C>
__global__ void __launch_bounds__(256) eigen_divide_zslab_kernel(C *__restrict__ A, const T *__restrict__ lx, const T *__restrict__ ly,
                                                                 const T *__restrict__ lz, int NxG, int Ny, int nz, int z0, int nlev, int zero_mode) {
    const long n = (long)NxG * Ny * nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % NxG), j = (int)((t / NxG) % Ny), k = (int)(t / ((long)NxG * Ny));
    C v = A[t];
    C o;
    if (z0 + k >= nlev) { o.x = 0; o.y = 0; A[t] = o; return; }   // zero padding levels of the half spectrum
    const T lam = lx[i] + ly[j] + lz[z0 + k];
    if (t == 0 && zero_mode) { o.x = 0; o.y = 0; }
    else { o.x = -v.x / lam; o.y = -v.y / lam; }
    A[t] = o;
}

// ϕ̂ = -b̂ / (λx + λy + λz) in the transposed layout; global mode (0,0,0) lives on rank 0
template <typename T, typename C>
__global__ void __launch_bounds__(256) eigen_divide_dist_kernel(C *__restrict__ A, const T *__restrict__ lx, const T *__restrict__ ly,
                                                                const T *__restrict__ lz, int NxG, int ny, int Nz, int y0, int zero_mode) {
    const long n = (long)NxG * ny * Nz;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % NxG), j = (int)((t / NxG) % ny), k = (int)(t / ((long)NxG * ny));
    C v = A[t];
    const T lam = lx[i] + ly[y0 + j] + lz[k];
    C o;
    if (t == 0 && zero_mode) { o.x = 0; o.y = 0; }
    else { o.x = -v.x / lam; o.y = -v.y / lam; }
    A[t] = o;
}

}  // namespace ob

template <typename T>
struct DistSolverT : ob_solver {
    using C = typename Cx<T>::type;
    int R = 1, rank = 0, nx = 0, ny = 0, nz = 0, NxG = 0;
    bool zx = false;   // x<->z transposition (FFT-based solvers with Nz % R == 0), else y<->x
    cufftHandle plan_xy = 0;
    bool has_xy = false;
    // zr: periodic z -> real-to-complex transform along z in the slab layout; the half spectrum (Nz/2+1 levels, padded
    // with zero levels to a multiple of R) is what gets transposed and transformed in (x, y): half the NVLink and HBM bytes
    bool zr = false;
    int Nzh = 0, NzT = 0;
    T *Rr = nullptr;
    cufftHandle plan_zr2c = 0, plan_zc2r = 0;
    bool use_ipc = false;           // pull transposes through CUDA-IPC peer mappings instead of NCCL send/recv
    ob::PeerPtrs<C> peerS, peerT;
    int *d_bar = nullptr;
    int *d_bflags = nullptr;   // [R] epochs published by the peers (IPC-exported)
    ob::PeerFlags peerF;
    int bar_epoch = 0;
    C *S = nullptr, *Tt = nullptr, *buf_a = nullptr, *buf_b = nullptr;
    // pipelined solve (IPC pulls): the local levels of the z-local layout are split into chunks; a second stream runs the pulls
    // (and the per-chunk barriers of the closing transposition) while the main stream transforms the chunks already there
    static constexpr int MAXCH = 8;
    bool pipelined = false;
    int nch = 0, ch0[MAXCH], chlen[MAXCH];
    cufftHandle plan_xy_ch[2] = {0, 0};   // plans for the two chunk lengths that occur
    int plan_xy_len[2] = {0, 0};
    cudaStream_t st2 = nullptr;
    int pull_ctas = 592;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_pull[MAXCH], ev_fft[MAXCH];
    // west-column mode (enable_west_column): the closing transposition writes rows of nx + 1 columns into S2 and the inverse z
    // transform runs on those (Rr2: its real output when zr)
    bool west = false;
    C *S2 = nullptr, *buf_a2 = nullptr;
    T *Rr2 = nullptr;
    cufftHandle plan_zc2r2 = 0, plan_z2 = 0;
    T *lam[3] = {nullptr, nullptr, nullptr};
    C *tw_f = nullptr, *tw_b = nullptr;          // z DCT twiddles (Bounded regular z)
    T *diag = nullptr, *lower = nullptr, *tscr = nullptr;
    cufftHandle plan_yz = 0, plan_y = 0, plan_z = 0, plan_x = 0;
    bool has_yz = false, has_y = false, has_z = false, has_x = false;
    double scale_ = 1.0;
    static constexpr cufftType CT = std::is_same<T, double>::value ? CUFFT_Z2Z : CUFFT_C2C;

    int32_t exec(cufftHandle p, C *data, int dir) {
        launches++;
        if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2Z(p, data, data, dir));
        else CUFFT_TRY(cufftExecC2C(p, data, data, dir));
        return OB_OK;
    }
    int32_t init(ob_ctx *c, const ob_grid_desc *g) {
        ctx = c;
        ft = g->float_type;
        R = c->world; rank = c->rank;
        for (int d = 0; d < 3; d++) { N[d] = g->N[d]; topo[d] = g->topology[d]; L[d] = g->L[d]; }
        nx = N[0]; NxG = nx * R;
        if (topo[0] != OB_PERIODIC || topo[1] != OB_PERIODIC) return fail(OB_ERR_UNSUPPORTED, "distributed solver: x and y must be Periodic");
        if (topo[2] == OB_FLAT) return fail(OB_ERR_UNSUPPORTED, "distributed solver: Flat z");
        tridiag = g->dzf_host != nullptr;
        zx = !tridiag && (N[2] % R == 0) && !getenv("OB_DIST_FORCE_YX");
        const bool zr_possible = !tridiag && g->topology[2] == OB_PERIODIC && N[2] > 1 && !getenv("OB_DIST_FORCE_YX") && !getenv("OB_SOLVER_NO_R2C");
        if (!zx && !zr_possible && N[1] % R) return fail(OB_ERR_INVALID, "distributed solver: Ny = %d (or Nz = %d) must be divisible by the number of ranks %d", N[1], N[2], R);
        zr = !tridiag && topo[2] == OB_PERIODIC && N[2] > 1 && !getenv("OB_DIST_FORCE_YX") && !getenv("OB_SOLVER_NO_R2C");
        if (zr) zx = true;
        Nzh = N[2] / 2 + 1;
        NzT = zr ? R * ((Nzh + R - 1) / R) : N[2];   // levels held by the complex slab storage
        ny = zx ? N[1] : N[1] / R;
        nz = zx ? NzT / R : N[2];
        const long n = (long)nx * N[1] * NzT;
        for (C **p : {&S, &Tt, &buf_a, &buf_b}) { CUDA_TRY(cudaMalloc(p, sizeof(C) * n)); CUDA_TRY(cudaMemsetAsync(*p, 0, sizeof(C) * n, ctx->stream)); }
        if (zr) {
            const long nr = (long)nx * N[1] * N[2];
            CUDA_TRY(cudaMalloc(&Rr, sizeof(T) * nr));
            CUDA_TRY(cudaMemsetAsync(Rr, 0, sizeof(T) * nr, ctx->stream));
        }
        // eigenvalues with the GLOBAL x extent (poisson_eigenvalues.jl:8-32)
        const int Ng[3] = {NxG, N[1], N[2]};
        const double Lg[3] = {L[0] * R, L[1], L[2]};
        for (int d = 0; d < 3; d++) {
            std::vector<T> h(Ng[d]);
            for (int i = 0; i < Ng[d]; i++) {
                double v = 0;
                if (topo[d] == OB_PERIODIC) { double s = 2 * sin(i * M_PI / Ng[d]) / (Lg[d] / Ng[d]); v = s * s; }
                else if (topo[d] == OB_BOUNDED) { double s = 2 * sin(i * M_PI / (2.0 * Ng[d])) / (Lg[d] / Ng[d]); v = s * s; }
                h[i] = (T)v;
            }
            CUDA_TRY(cudaMalloc(&lam[d], sizeof(T) * Ng[d]));
            CUDA_TRY(cudaMemcpy(lam[d], h.data(), sizeof(T) * Ng[d], cudaMemcpyHostToDevice));
        }
        scale_ = 1.0 / ((double)NxG * N[1]);
        const bool z_fft = !tridiag;
        if (z_fft) scale_ /= N[2];
        if (zx) {
            int nzz[1] = {N[2]};   // z in the slab layout: stride nx*Ny, contiguous batch nx*Ny
            if (zr) {
                constexpr cufftType FWD = std::is_same<T, double>::value ? CUFFT_D2Z : CUFFT_R2C;
                constexpr cufftType BWD = std::is_same<T, double>::value ? CUFFT_Z2D : CUFFT_C2R;
                int nre[1] = {N[2]}, nco[1] = {Nzh};
                CUFFT_TRY(cufftPlanMany(&plan_zr2c, 1, nzz, nre, nx * N[1], 1, nco, nx * N[1], 1, FWD, nx * N[1]));
                CUFFT_TRY(cufftPlanMany(&plan_zc2r, 1, nzz, nco, nx * N[1], 1, nre, nx * N[1], 1, BWD, nx * N[1]));
                CUFFT_TRY(cufftSetStream(plan_zr2c, ctx->stream));
                CUFFT_TRY(cufftSetStream(plan_zc2r, ctx->stream));
            } else {
                CUFFT_TRY(cufftPlanMany(&plan_z, 1, nzz, nzz, nx * N[1], 1, nzz, nx * N[1], 1, CT, nx * N[1]));
                CUFFT_TRY(cufftSetStream(plan_z, ctx->stream));
                has_z = true;
            }
            int nn[2] = {N[1], NxG};   // (x, y) in the z-local layout: contiguous, batched over the local levels
            CUFFT_TRY(cufftPlanMany(&plan_xy, 2, nn, nullptr, 1, NxG * N[1], nullptr, 1, NxG * N[1], CT, nz));
            CUFFT_TRY(cufftSetStream(plan_xy, ctx->stream));
            has_xy = true;
            if (topo[2] == OB_BOUNDED) {
                std::vector<C> f(N[2]), b(N[2]);
                for (int k = 0; k < N[2]; k++) {
                    double a = -2 * M_PI * k / (4.0 * N[2]);
                    f[k].x = (T)cos(a); f[k].y = (T)sin(a); b[k].x = (T)cos(-a); b[k].y = (T)sin(-a);
                }
                b[0].x *= (T)0.5; b[0].y *= (T)0.5;
                CUDA_TRY(cudaMalloc(&tw_f, sizeof(C) * N[2])); CUDA_TRY(cudaMalloc(&tw_b, sizeof(C) * N[2]));
                CUDA_TRY(cudaMemcpy(tw_f, f.data(), sizeof(C) * N[2], cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(tw_b, b.data(), sizeof(C) * N[2], cudaMemcpyHostToDevice));
            }
        } else if (z_fft && topo[2] == OB_PERIODIC) {   // one strided rank-2 (z, y) transform per x column
            int nn[2] = {N[2], N[1]};
            CUFFT_TRY(cufftPlanMany(&plan_yz, 2, nn, nn, nx, 1, nn, nx, 1, CT, nx));
            CUFFT_TRY(cufftSetStream(plan_yz, ctx->stream));
            has_yz = true;
        } else {
            int nn[1] = {N[1]};  // y: one z-plane per call (stride nx, batch nx)
            CUFFT_TRY(cufftPlanMany(&plan_y, 1, nn, nn, nx, 1, nn, nx, 1, CT, nx));
            CUFFT_TRY(cufftSetStream(plan_y, ctx->stream));
            has_y = true;
            if (z_fft) {  // Bounded regular z: Makhoul DCT through a z-FFT
                int nz[1] = {N[2]};
                CUFFT_TRY(cufftPlanMany(&plan_z, 1, nz, nz, nx * N[1], 1, nz, nx * N[1], 1, CT, nx * N[1]));
                CUFFT_TRY(cufftSetStream(plan_z, ctx->stream));
                has_z = true;
                std::vector<C> f(N[2]), b(N[2]);
                for (int k = 0; k < N[2]; k++) {
                    double a = -2 * M_PI * k / (4.0 * N[2]);
                    f[k].x = (T)cos(a); f[k].y = (T)sin(a); b[k].x = (T)cos(-a); b[k].y = (T)sin(-a);
                }
                b[0].x *= (T)0.5; b[0].y *= (T)0.5;
                CUDA_TRY(cudaMalloc(&tw_f, sizeof(C) * N[2])); CUDA_TRY(cudaMalloc(&tw_b, sizeof(C) * N[2]));
                CUDA_TRY(cudaMemcpy(tw_f, f.data(), sizeof(C) * N[2], cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(tw_b, b.data(), sizeof(C) * N[2], cudaMemcpyHostToDevice));
            }
        }
        if (!zx) {
            int nn[1] = {NxG};
            CUFFT_TRY(cufftPlanMany(&plan_x, 1, nn, nn, 1, NxG, nn, 1, NxG, CT, ny * N[2]));
            CUFFT_TRY(cufftSetStream(plan_x, ctx->stream));
            has_x = true;
        }
        if (zx && !getenv("OB_DIST_NO_IPC")) OB_TRY(setup_ipc());
        // pipelined transposes pay where the pulls are NVLink-bound (7/8 of a pull is remote on 8 ranks); on 2 ranks they are
        // HBM-bound like the transforms they would hide behind (measured: +0.03 ms per solve), so the default is R >= 4
        const char *pe = getenv("OB_DIST_PIPELINE");
        const bool want_pipe = pe ? atoi(pe) != 0 : (R >= 4 && !getenv("OB_DIST_NO_PIPELINE"));
        if (zx && use_ipc && nz >= 2 && R > 1 && want_pipe) OB_TRY(setup_pipeline());
        if (tridiag) {  // diagonal in the transposed layout (fourier_tridiagonal_poisson_solver.jl:199-229)
            const int Nz = N[2], Hz = g->H[2];
            const T *dzf = (const T *)g->dzf_host, *dzc = (const T *)g->dzc_host;
            auto DZF = [&](int k) { return dzf[k + Hz]; };
            auto DZC = [&](int k) { return dzc[k + Hz - 1]; };
            std::vector<T> lx(NxG), ly(N[1]);
            CUDA_TRY(cudaMemcpy(lx.data(), lam[0], sizeof(T) * NxG, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(ly.data(), lam[1], sizeof(T) * N[1], cudaMemcpyDeviceToHost));
            std::vector<T> D((size_t)n), low(std::max(1, Nz - 1));
            for (int k = 1; k <= Nz; k++)
                for (int j = 0; j < ny; j++)
                    for (int i = 0; i < NxG; i++) {
                        T l = lx[i] + ly[rank * ny + j];
                        T v;
                        if (k == 1) v = (T)-1 / DZF(2) - DZC(1) * l;
                        else if (k == Nz) v = (T)-1 / DZF(Nz) - DZC(Nz) * l;
                        else v = -((T)1 / DZF(k + 1) + (T)1 / DZF(k)) - DZC(k) * l;
                        D[i + (size_t)NxG * (j + (size_t)ny * (k - 1))] = v;
                    }
            for (int q = 1; q <= Nz - 1; q++) low[q - 1] = (T)1 / DZF(q + 1);
            CUDA_TRY(cudaMalloc(&diag, sizeof(T) * n)); CUDA_TRY(cudaMalloc(&tscr, sizeof(T) * n));
            CUDA_TRY(cudaMalloc(&lower, sizeof(T) * low.size()));
            CUDA_TRY(cudaMemcpy(diag, D.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(lower, low.data(), sizeof(T) * low.size(), cudaMemcpyHostToDevice));
        }
        return OB_OK;
    }
    // exchange cudaIpcMemHandle_t of S and Tt between the ranks (all-gather over NCCL) and map the peers' arrays
    int32_t setup_ipc() {
        if (R > OB_MAX_PEERS) return OB_OK;
        struct Pair { cudaIpcMemHandle_t s, t, f; };
        Pair mine;
        CUDA_TRY(cudaMalloc(&d_bflags, sizeof(int) * OB_MAX_PEERS));
        CUDA_TRY(cudaMemsetAsync(d_bflags, 0, sizeof(int) * OB_MAX_PEERS, ctx->stream));
        if (cudaIpcGetMemHandle(&mine.s, S) != cudaSuccess || cudaIpcGetMemHandle(&mine.t, Tt) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.f, d_bflags) != cudaSuccess) { cudaGetLastError(); return OB_OK; }
        Pair *d_all = nullptr;
        CUDA_TRY(cudaMalloc(&d_all, sizeof(Pair) * R));
        CUDA_TRY(cudaMemcpyAsync(d_all + rank, &mine, sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ncclAllGather(d_all + rank, d_all, sizeof(Pair), ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
        std::vector<Pair> all(R);
        CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(Pair) * R, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_all);
        int ok = 1;
        for (int r = 0; r < R; r++) {
            if (r == rank) { peerS.p[r] = S; peerT.p[r] = Tt; peerF.p[r] = d_bflags; continue; }
            void *ps = nullptr, *pt = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&ps, all[r].s, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&pt, all[r].t, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&pf, all[r].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            peerS.p[r] = (const C *)ps; peerT.p[r] = (const C *)pt; peerF.p[r] = (int *)pf;
        }
        // every rank must take the same path: agree through an all-reduce (min)
        CUDA_TRY(cudaMalloc(&d_bar, sizeof(int)));
        CUDA_TRY(cudaMemcpyAsync(d_bar, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ncclAllReduce(d_bar, d_bar, 1, ncclInt, ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&ok, d_bar, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        use_ipc = ok != 0;
        return OB_OK;
    }
    int32_t setup_pipeline() {
        const char *ce = getenv("OB_DIST_CHUNKS");   // tuning knob; measured at N = 8, 256^3 per GPU: 2 chunks 9.78 ms/step, 4: 9.90, 8: 10.24
        nch = std::max(1, std::min(std::min((int)MAXCH, nz), ce ? atoi(ce) : 2));
        int z = 0;
        for (int c = 0; c < nch; c++) { chlen[c] = nz / nch + (c < nz % nch ? 1 : 0); ch0[c] = z; z += chlen[c]; }
        int nn[2] = {N[1], NxG};
        for (int c = 0; c < nch; c++) {
            int slot = plan_xy_len[0] == chlen[c] ? 0 : plan_xy_len[1] == chlen[c] ? 1 : plan_xy_len[0] == 0 ? 0 : 1;
            if (plan_xy_len[slot] == chlen[c]) continue;
            CUFFT_TRY(cufftPlanMany(&plan_xy_ch[slot], 2, nn, nullptr, 1, NxG * N[1], nullptr, 1, NxG * N[1], CT, chlen[c]));
            CUFFT_TRY(cufftSetStream(plan_xy_ch[slot], ctx->stream));
            plan_xy_len[slot] = chlen[c];
        }
        // the pulls run at the lowest priority (the context's stream has the highest: ob_init), so the transforms of the main
        // stream get the CTA slots they ask for and the pulls fill what is left
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, least));
        pull_ctas = (ctx->sm_count > 0 ? ctx->sm_count : 148) * 4;
        CUDA_TRY(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
        for (int c = 0; c < nch; c++) {
            CUDA_TRY(cudaEventCreateWithFlags(&ev_pull[c], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_fft[c], cudaEventDisableTiming));
        }
        pipelined = true;
        return OB_OK;
    }
    // forward pull, (x, y) transforms + eigenvalue division, closing pull -- chunk by chunk on two streams.  Every rank issues
    // the same sequence of barriers (one on the main stream, then one per chunk on the second stream), so the epochs agree.
    int32_t pipelined_middle() {
        cudaStream_t st = ctx->stream;
        const long plane = (long)NxG * N[1];
        OB_TRY(barrier());        // every peer's z transform is complete
        CUDA_TRY(cudaEventRecord(ev_ready, st));
        CUDA_TRY(cudaStreamWaitEvent(st2, ev_ready, 0));
        for (int c = 0; c < nch; c++) {
            pull_z_to_x_chunk_kernel<C><<<std::min((unsigned)pull_ctas, nblk(plane * chlen[c], 1024)), 256, 0, st2>>>(peerS, Tt, nx, NxG, N[1], nz, rank, ch0[c], chlen[c]);
            CUDA_TRY(cudaEventRecord(ev_pull[c], st2));
        }
        const int nxo = nx + (west ? 1 : 0);
        // (never S: a peer may still be pulling this rank's z-transformed slab out of it when the first chunk comes back)
        C *out = west ? S2 : buf_b;
        for (int c = 0; c < nch; c++) {
            CUDA_TRY(cudaStreamWaitEvent(st, ev_pull[c], 0));
            C *chunk = Tt + plane * ch0[c];
            cufftHandle pl = plan_xy_len[0] == chlen[c] ? plan_xy_ch[0] : plan_xy_ch[1];
            OB_TRY(exec(pl, chunk, CUFFT_FORWARD));
            eigen_divide_zslab_kernel<T, C><<<nblk(plane * chlen[c], 256), 256, 0, st>>>(chunk, lam[0], lam[1], lam[2], NxG, N[1], chlen[c], rank * nz + ch0[c],
                                                                                         zr ? Nzh : N[2], (rank == 0 && c == 0) ? 1 : 0);
            OB_TRY(exec(pl, chunk, CUFFT_INVERSE));
            CUDA_TRY(cudaEventRecord(ev_fft[c], st));
            // closing transposition of this chunk: every peer's chunk c must be transformed
            CUDA_TRY(cudaStreamWaitEvent(st2, ev_fft[c], 0));
            ipc_barrier_kernel<<<1, 32, 0, st2>>>(peerF, d_bflags, R, rank, ++bar_epoch);
            pull_x_to_z_chunk_kernel<C><<<std::min((unsigned)pull_ctas, nblk((long)nxo * N[1] * chlen[c] * R, 1024)), 256, 0, st2>>>(peerT, out, nx, NxG, N[1], nz, R, rank, west ? 1 : 0,
                                                                                                      ch0[c], chlen[c]);
            launches += 4;
        }
        CUDA_TRY(cudaEventRecord(ev_done, st2));
        CUDA_TRY(cudaStreamWaitEvent(st, ev_done, 0));
        launches += nch;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // inverse z transform of the slab layout, reading the transposed spectrum from `in`; rows of nx columns (result in S / Rr) or
    // nx + 1 columns (`in` = S2, result in S2 / Rr2)
    int32_t inverse_z(C *in, bool wide) {
        cudaStream_t st = ctx->stream;
        const bool z_dct = !tridiag && topo[2] == OB_BOUNDED;
        const int nxo = nx + (wide ? 1 : 0);
        const unsigned nbo = nblk((long)nxo * N[1] * NzT, 256);
        C *dst = wide ? S2 : S, *tmp = wide ? buf_a2 : buf_a;
        if (z_dct) {
            twiddle_bwd_kernel<T, C><<<nbo, 256, 0, st>>>(in, tmp, tw_b, nxo, N[1], N[2], 2);
            OB_TRY(exec(wide ? plan_z2 : plan_z, tmp, CUFFT_INVERSE));
            unpermute_kernel<C><<<nbo, 256, 0, st>>>(tmp, dst, nxo, N[1], N[2], 2);
            launches += 2;
        } else if (zr) {
            launches++;
            if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2D(wide ? plan_zc2r2 : plan_zc2r, in, wide ? Rr2 : Rr));
            else CUFFT_TRY(cufftExecC2R(wide ? plan_zc2r2 : plan_zc2r, in, wide ? Rr2 : Rr));
        } else {
            launches++;
            if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2Z(wide ? plan_z2 : plan_z, in, dst, CUFFT_INVERSE));
            else CUFFT_TRY(cufftExecC2C(wide ? plan_z2 : plan_z, in, dst, CUFFT_INVERSE));
        }
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    bool enable_west_column() override {
        if (west) return true;
        if (!zx || !use_ipc || getenv("OB_DIST_NO_WEST_COLUMN")) return false;
        const int nx1 = nx + 1;
        const long n1 = (long)nx1 * N[1] * NzT;
        if (cudaMalloc(&S2, sizeof(C) * n1) != cudaSuccess) { cudaGetLastError(); return false; }
        cudaMemsetAsync(S2, 0, sizeof(C) * n1, ctx->stream);
        int nzz[1] = {N[2]};
        bool ok = true;
        if (zr) {
            constexpr cufftType BWD = std::is_same<T, double>::value ? CUFFT_Z2D : CUFFT_C2R;
            int nre[1] = {N[2]}, nco[1] = {Nzh};
            ok = cudaMalloc(&Rr2, sizeof(T) * (long)nx1 * N[1] * N[2]) == cudaSuccess &&
                 cufftPlanMany(&plan_zc2r2, 1, nzz, nco, nx1 * N[1], 1, nre, nx1 * N[1], 1, BWD, nx1 * N[1]) == CUFFT_SUCCESS &&
                 cufftSetStream(plan_zc2r2, ctx->stream) == CUFFT_SUCCESS;
        } else {
            ok = cufftPlanMany(&plan_z2, 1, nzz, nzz, nx1 * N[1], 1, nzz, nx1 * N[1], 1, CT, nx1 * N[1]) == CUFFT_SUCCESS &&
                 cufftSetStream(plan_z2, ctx->stream) == CUFFT_SUCCESS;
            if (ok && topo[2] == OB_BOUNDED) ok = cudaMalloc(&buf_a2, sizeof(C) * n1) == cudaSuccess;
        }
        if (!ok) { cudaGetLastError(); return false; }
        west = true;
        return true;
    }
    const void *solution() override {
        if (!west) return storage();
        return zr ? (const void *)(Rr2 + 1) : (const void *)(S2 + 1);
    }
    long solution_ldx() const override { return west ? nx + 1 : nx; }
    bool has_west_column() const override { return west; }
    // stream-ordered barrier across the ranks
    int32_t barrier() {
        ipc_barrier_kernel<<<1, 32, 0, ctx->stream>>>(peerF, d_bflags, R, rank, ++bar_epoch);
        launches++;
        return OB_OK;
    }
    ~DistSolverT() override {
        if (use_ipc) {
            cudaStreamSynchronize(ctx->stream);
            for (int r = 0; r < R; r++) if (r != rank) { cudaIpcCloseMemHandle((void *)peerS.p[r]); cudaIpcCloseMemHandle((void *)peerT.p[r]); cudaIpcCloseMemHandle((void *)peerF.p[r]); }
        }
        cudaFree(d_bar); cudaFree(d_bflags);
        cudaFree(S); cudaFree(Tt); cudaFree(buf_a); cudaFree(buf_b);
        for (int d = 0; d < 3; d++) cudaFree(lam[d]);
        cudaFree(tw_f); cudaFree(tw_b); cudaFree(diag); cudaFree(lower); cudaFree(tscr);
        if (has_yz) cufftDestroy(plan_yz);
        if (has_y) cufftDestroy(plan_y);
        if (has_z) cufftDestroy(plan_z);
        if (has_x) cufftDestroy(plan_x);
        if (has_xy) cufftDestroy(plan_xy);
        if (zr) { cufftDestroy(plan_zr2c); cufftDestroy(plan_zc2r); cudaFree(Rr); }
        if (pipelined) {
            cudaStreamSynchronize(st2); cudaStreamDestroy(st2);
            cudaEventDestroy(ev_ready); cudaEventDestroy(ev_done);
            for (int c = 0; c < nch; c++) { cudaEventDestroy(ev_pull[c]); cudaEventDestroy(ev_fft[c]); }
            for (int q = 0; q < 2; q++) if (plan_xy_ch[q]) cufftDestroy(plan_xy_ch[q]);
        }
        if (plan_zc2r2) cufftDestroy(plan_zc2r2);
        if (plan_z2) cufftDestroy(plan_z2);
        cudaFree(S2); cudaFree(buf_a2); cudaFree(Rr2);
    }
    void *storage() override { return zr ? (void *)Rr : (void *)S; }
    double scale() override { return scale_; }
    bool real_storage() const override { return zr; }

    // all-to-all of per-peer chunks (nccl_transpose.jl:47-75: grouped Send/Recv, complex as 2 x real)
    int32_t alltoall(const C *send, C *recv) {
        const size_t chunk = (size_t)nx * ny * nz;
        ncclComm_t comm = (ncclComm_t)ctx->comm;
        const ncclDataType_t dt = std::is_same<T, double>::value ? ncclDouble : ncclFloat;
        // the self chunk is a local copy (nccl_transpose.jl:47-75)
        CUDA_TRY(cudaMemcpyAsync(recv + rank * chunk, send + rank * chunk, sizeof(C) * chunk, cudaMemcpyDeviceToDevice, ctx->stream));
        NCCL_TRY(ncclGroupStart());
        for (int r = 0; r < R; r++) {
            if (r == rank) continue;
            NCCL_TRY(ncclSend(send + r * chunk, 2 * chunk, dt, r, comm, ctx->stream));
            NCCL_TRY(ncclRecv(recv + r * chunk, 2 * chunk, dt, r, comm, ctx->stream));
        }
        NCCL_TRY(ncclGroupEnd());
        launches++;
        return OB_OK;
    }
    int32_t y_fft(int dir) {
        for (int k = 0; k < N[2]; k++) OB_TRY(exec(plan_y, S + (long)k * nx * N[1], dir));
        return OB_OK;
    }
    int32_t solve_in_storage() override {
        const long n = (long)nx * N[1] * NzT;
        const unsigned nb = nblk(n, 256);
        cudaStream_t st = ctx->stream;
        const bool z_dct = !tridiag && topo[2] == OB_BOUNDED;
        if (zx) {
            // z transform in the slab layout (DCT through permute / FFT / twiddle when Bounded)
            if (z_dct) {
                permute_kernel<C><<<nb, 256, 0, st>>>(S, buf_a, nx, N[1], N[2], 2);
                OB_TRY(exec(plan_z, buf_a, CUFFT_FORWARD));
                twiddle_fwd_kernel<T, C><<<nb, 256, 0, st>>>(buf_a, S, tw_f, nx, N[1], N[2], 2);
                launches += 2;
            } else if (zr) {
                launches++;
                if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecD2Z(plan_zr2c, Rr, S));
                else CUFFT_TRY(cufftExecR2C(plan_zr2c, Rr, S));
            } else {
                OB_TRY(exec(plan_z, S, CUFFT_FORWARD));
            }
            if (pipelined) {
                OB_TRY(pipelined_middle());
                return inverse_z(west ? S2 : buf_b, west);
            }
            if (use_ipc) {
                OB_TRY(barrier());        // every peer's z transform is complete
                pull_z_to_x_kernel<C><<<nb, 256, 0, st>>>(peerS, Tt, nx, NxG, N[1], nz, NzT, rank);
            } else {
                OB_TRY(alltoall(S, buf_b));   // chunks are contiguous in S: no pack
                unpack_z_to_x_kernel<C><<<nb, 256, 0, st>>>(buf_b, Tt, nx, NxG, N[1], nz);
            }
            OB_TRY(exec(plan_xy, Tt, CUFFT_FORWARD));
            eigen_divide_zslab_kernel<T, C><<<nb, 256, 0, st>>>(Tt, lam[0], lam[1], lam[2], NxG, N[1], nz, rank * nz, zr ? Nzh : N[2], rank == 0 ? 1 : 0);
            OB_TRY(exec(plan_xy, Tt, CUFFT_INVERSE));
            if (west) {
                // closing transposition with the west neighbour's last column in front of every row, inverse z transform on
                // rows of nx + 1 columns
                const int nx1 = nx + 1;
                const long n1 = (long)nx1 * N[1] * NzT;
                const unsigned nb1 = nblk(n1, 256);
                OB_TRY(barrier());
                pull_x_to_z_west_kernel<C><<<nb1, 256, 0, st>>>(peerT, S2, nx, NxG, N[1], nz, NzT, rank);
                launches += 3;
                if (z_dct) {
                    twiddle_bwd_kernel<T, C><<<nb1, 256, 0, st>>>(S2, buf_a2, tw_b, nx1, N[1], N[2], 2);
                    OB_TRY(exec(plan_z2, buf_a2, CUFFT_INVERSE));
                    unpermute_kernel<C><<<nb1, 256, 0, st>>>(buf_a2, S2, nx1, N[1], N[2], 2);
                    launches += 2;
                } else if (zr) {
                    launches++;
                    if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2D(plan_zc2r2, S2, Rr2));
                    else CUFFT_TRY(cufftExecC2R(plan_zc2r2, S2, Rr2));
                } else {
                    OB_TRY(exec(plan_z2, S2, CUFFT_INVERSE));
                }
                CUDA_TRY(cudaGetLastError());
                return OB_OK;
            }
            if (use_ipc) {
                OB_TRY(barrier());        // every peer's inverse (x, y) transform is complete (and its forward pull long done)
                pull_x_to_z_kernel<C><<<nb, 256, 0, st>>>(peerT, S, nx, NxG, N[1], nz, NzT, rank);
            } else {
                pack_x_to_z_kernel<C><<<nb, 256, 0, st>>>(Tt, buf_a, nx, NxG, N[1], nz);
                OB_TRY(alltoall(buf_a, S));   // received chunks land contiguously in S: no unpack
            }
            launches += 3;
            if (z_dct) {
                twiddle_bwd_kernel<T, C><<<nb, 256, 0, st>>>(S, buf_a, tw_b, nx, N[1], N[2], 2);
                OB_TRY(exec(plan_z, buf_a, CUFFT_INVERSE));
                unpermute_kernel<C><<<nb, 256, 0, st>>>(buf_a, S, nx, N[1], N[2], 2);
                launches += 2;
            } else if (zr) {
                launches++;
                if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2D(plan_zc2r, S, Rr));
                else CUFFT_TRY(cufftExecC2R(plan_zc2r, S, Rr));
            } else {
                OB_TRY(exec(plan_z, S, CUFFT_INVERSE));
            }
            CUDA_TRY(cudaGetLastError());
            return OB_OK;
        }
        // forward: z (Bounded first), y in the slab layout
        if (z_dct) {
            permute_kernel<C><<<nb, 256, 0, st>>>(S, buf_a, nx, N[1], N[2], 2);
            OB_TRY(exec(plan_z, buf_a, CUFFT_FORWARD));
            twiddle_fwd_kernel<T, C><<<nb, 256, 0, st>>>(buf_a, S, tw_f, nx, N[1], N[2], 2);
            launches += 2;
        }
        if (has_yz) OB_TRY(exec(plan_yz, S, CUFFT_FORWARD));
        else OB_TRY(y_fft(CUFFT_FORWARD));
        // y -> x transpose, x transform
        pack_y_to_x_kernel<C><<<nb, 256, 0, st>>>(S, buf_a, nx, ny, N[1], N[2]);
        OB_TRY(alltoall(buf_a, buf_b));
        unpack_y_to_x_kernel<C><<<nb, 256, 0, st>>>(buf_b, Tt, nx, ny, NxG, N[2]);
        OB_TRY(exec(plan_x, Tt, CUFFT_FORWARD));
        if (tridiag) {
            dim3 grid(nblk(NxG, 128), ny);
            thomas_kernel<T, C><<<grid, 128, 0, st>>>(Tt, lower, lower, diag, tscr, NxG, ny, N[2], (T)(10 * std::numeric_limits<T>::epsilon()), rank == 0 ? 1 : 0);
        } else {
            eigen_divide_dist_kernel<T, C><<<nb, 256, 0, st>>>(Tt, lam[0], lam[1], lam[2], NxG, ny, N[2], rank * ny, rank == 0 ? 1 : 0);
        }
        OB_TRY(exec(plan_x, Tt, CUFFT_INVERSE));
        pack_x_to_y_kernel<C><<<nb, 256, 0, st>>>(Tt, buf_a, nx, ny, NxG, N[2]);
        OB_TRY(alltoall(buf_a, buf_b));
        unpack_x_to_y_kernel<C><<<nb, 256, 0, st>>>(buf_b, S, nx, ny, N[1], N[2]);
        launches += 5;
        if (has_yz) OB_TRY(exec(plan_yz, S, CUFFT_INVERSE));
        else OB_TRY(y_fft(CUFFT_INVERSE));
        if (z_dct) {
            twiddle_bwd_kernel<T, C><<<nb, 256, 0, st>>>(S, buf_a, tw_b, nx, N[1], N[2], 2);
            OB_TRY(exec(plan_z, buf_a, CUFFT_INVERSE));
            unpermute_kernel<C><<<nb, 256, 0, st>>>(buf_a, S, nx, N[1], N[2], 2);
            launches += 2;
        }
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
};
