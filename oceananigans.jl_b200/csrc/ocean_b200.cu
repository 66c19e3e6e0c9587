// ocean_b200.cu -- libocean_b200.so: host orchestration + C ABI (include/ocean_b200.h).
//
// The time-step driver mirrors the reference call stack:
//   time_step! RK3/AB2     src/TimeSteppers/runge_kutta_3.jl:103-168, quasi_adams_bashforth_2.jl:90-126
//   rk3_substep!/ab2_step! src/Models/NonhydrostaticModels/nonhydrostatic_rk3_substep.jl:31-63, nonhydrostatic_ab2_step.jl:10-57
//   pressure correction    .../pressure_correction.jl:6-106, solve_for_pressure.jl:12-126
//   update_state!          .../update_nonhydrostatic_model_state.jl:22-83
// but every step is a hand-written sm_100a kernel (or a cuFFT batched transform), batched over fields.
// There is no CPU fallback anywhere in this file.
#include "../../include/ocean_b200.h"

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <initializer_list>
#include <cmath>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "streaming.cuh"
#include "poisson.cuh"
#include "stencils.cuh"
#include "tendency.cuh"
#include "closures.cuh"
#include "dynsmag.cuh"
#include "tend_launch.h"

using namespace ob;

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int32_t fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(x)                                                                                      \
    do {                                                                                                 \
        cudaError_t e_ = (x);                                                                            \
        if (e_ != cudaSuccess) return fail(OB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)
#define CUFFT_TRY(x)                                                                                     \
    do {                                                                                                 \
        cufftResult r_ = (x);                                                                            \
        if (r_ != CUFFT_SUCCESS) return fail(OB_ERR_CUFFT, "%s:%d %s: cufft error %d", __FILE__, __LINE__, #x, (int)r_); \
    } while (0)
#define OB_TRY(x)              \
    do {                       \
        int32_t s_ = (x);      \
        if (s_ != OB_OK) return s_; \
    } while (0)

extern "C" const char *ob_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
struct ob_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int *d_flag = nullptr;
    int rank = 0, world = 1;
    void *comm = nullptr;  // ncclComm_t (dist.cuh)
    cudaEvent_t tA = nullptr, tB = nullptr;
};

extern "C" int32_t ob_device_count(int32_t *n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return fail(OB_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n = c;
    return OB_OK;
}

extern "C" int32_t ob_init(int32_t device, ob_ctx **out) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess || c == 0)
        return fail(OB_ERR_NO_DEVICE, "no CUDA device: libocean_b200 has no CPU fallback");
    if (device < 0 || device >= c) return fail(OB_ERR_INVALID, "device %d out of range (%d devices)", device, c);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(OB_ERR_NO_DEVICE, "device %d is sm_%d%d; libocean_b200 is built for sm_100a only", device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    ob_ctx *ctx = new ob_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    {   // highest priority: helper streams of the library (distributed transposes) run below it
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, greatest));
    }
    CUDA_TRY(cudaMalloc(&ctx->d_flag, sizeof(int)));
    *out = ctx;
    return OB_OK;
}
extern "C" int32_t ob_shutdown(ob_ctx *ctx) {
    if (!ctx) return OB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_flag);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return OB_OK;
}
// device-side timer on the context's stream (bench.py: CUDA events on the stream the kernels are launched on)
extern "C" int32_t ob_timer_start(ob_ctx *ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->tA) { CUDA_TRY(cudaEventCreate(&ctx->tA)); CUDA_TRY(cudaEventCreate(&ctx->tB)); }
    CUDA_TRY(cudaEventRecord(ctx->tA, ctx->stream));
    return OB_OK;
}
extern "C" int32_t ob_timer_stop(ob_ctx *ctx, double *ms) {
    if (!ctx->tA) return fail(OB_ERR_INVALID, "ob_timer_stop without ob_timer_start");
    CUDA_TRY(cudaEventRecord(ctx->tB, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->tB));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->tA, ctx->tB));
    *ms = f;
    return OB_OK;
}
// FP64 issue-rate microbenchmark (bench.py: the second roofline of the FP64-bound tendency kernel): 8 independent DFMA
// chains per thread, enough CTAs to fill the device; returns thread-level FP64 instructions per second.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}
extern "C" int32_t ob_fp64_peak(ob_ctx *ctx, double *instr_per_s) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double)));
    const int iters = 20000, blocks = ctx->sm_count * 8;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
    fp64_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d, 100);
    CUDA_TRY(cudaEventRecord(a, ctx->stream));
    fp64_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d, iters);
    CUDA_TRY(cudaEventRecord(b, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    *instr_per_s = (double)blocks * 256.0 * iters * 8.0 / (ms * 1e-3);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return OB_OK;
}
extern "C" int32_t ob_sync(ob_ctx *ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OB_OK;
}
extern "C" int32_t ob_malloc(ob_ctx *ctx, size_t bytes, void **ptr) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 1));
    CUDA_TRY(cudaMemsetAsync(*ptr, 0, bytes, ctx->stream));
    return OB_OK;
}
extern "C" int32_t ob_free(ob_ctx *ctx, void *ptr) {
    (void)ctx;
    if (ptr) CUDA_TRY(cudaFree(ptr));
    return OB_OK;
}
extern "C" int32_t ob_malloc_host(ob_ctx *ctx, size_t bytes, void **ptr) {
    (void)ctx;
    CUDA_TRY(cudaMallocHost(ptr, bytes ? bytes : 1));
    return OB_OK;
}
extern "C" int32_t ob_free_host(ob_ctx *ctx, void *ptr) {
    (void)ctx;
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return OB_OK;
}
extern "C" int32_t ob_memcpy_h2d(ob_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return fail(OB_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));   // the caller may have made another device current (multi-context hosts)
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return OB_OK;
}
extern "C" int32_t ob_memcpy_d2h(ob_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return fail(OB_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return OB_OK;
}
// stream-ordered device -> pinned-host copy: the host may read dst after ob_sync (or after a later blocking call on ctx)
extern "C" int32_t ob_memcpy_d2h_async(ob_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return fail(OB_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));   // the caller may have made another device current (multi-context hosts)
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return OB_OK;
}
// make every later call on `ctx` wait for the work submitted so far on `other` (cudaStreamWaitEvent across contexts of
// the same device): joins the lanes of a host-streamed ensemble (streaming.py) without blocking the host
extern "C" int32_t ob_stream_wait(ob_ctx *ctx, ob_ctx *other) {
    if (!ctx || !other) return fail(OB_ERR_INVALID, "null context");
    if (ctx->device != other->device) return fail(OB_ERR_INVALID, "ob_stream_wait: contexts on different devices");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ev, other->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev, 0));
    CUDA_TRY(cudaEventDestroy(ev));   // released once the wait has been satisfied
    return OB_OK;
}
extern "C" int32_t ob_memcpy_d2d(ob_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return fail(OB_ERR_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(ctx->device));   // the caller may have made another device current (multi-context hosts)
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return OB_OK;
}
static inline unsigned nblk(long n, int b) { return (unsigned)((n + b - 1) / b); }
extern "C" int32_t ob_fill(ob_ctx *ctx, void *ptr, size_t n, int32_t ft, double value) {
    if (n == 0) return OB_OK;
    if (ft == OB_F64) fill_kernel<double><<<nblk(n, 256), 256, 0, ctx->stream>>>((double *)ptr, (long)n, value);
    else fill_kernel<float><<<nblk(n, 256), 256, 0, ctx->stream>>>((float *)ptr, (long)n, (float)value);
    CUDA_TRY(cudaGetLastError());
    return OB_OK;
}
extern "C" int32_t ob_any_nan(ob_ctx *ctx, const void *ptr, size_t n, int32_t ft, int32_t *flag) {
    CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    unsigned blocks = std::min<unsigned>(nblk(n, 256), 148 * 8);
    if (n) {
        if (ft == OB_F64) any_nan_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double *)ptr, (long)n, ctx->d_flag);
        else any_nan_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float *)ptr, (long)n, ctx->d_flag);
    }
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *flag = h;
    return OB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Poisson solver (host side)
// ---------------------------------------------------------------------------------------------------------------
struct ob_solver {
    ob_ctx *ctx = nullptr;
    int ft = OB_F64;
    int N[3] = {1, 1, 1}, topo[3] = {0, 0, 0};
    double L[3] = {1, 1, 1};
    bool tridiag = false;
    int64_t launches = 0;
    virtual ~ob_solver() {}
    // rhs has been written (as complex numbers) into storage(); result is left in storage(), to be scaled by `scale()`
    virtual int32_t solve_in_storage() = 0;
    virtual void *storage() = 0;
    virtual double scale() = 0;
    // true: storage() holds REAL numbers (Nx,Ny,Nz) on input and output (real-to-complex transforms inside)
    virtual bool real_storage() const { return false; }
    // true: level k of storage() lives at the Makhoul-permuted position (real DCT path)
    virtual bool z_permuted() const { return false; }
    // Distributed solvers can leave the solution with ONE EXTRA WEST COLUMN (the last column of the west neighbour's slab,
    // taken along by the closing transposition), so that the projection can form dp/dx at i = 1 without a halo exchange of
    // the pressure.  Off unless enabled; solution() then points at column i = 1 of rows of solution_ldx() elements (column
    // i = 0 is one element -- one complex number for complex storage -- in front of every row).
    virtual bool enable_west_column() { return false; }
    virtual const void *solution() { return storage(); }
    virtual long solution_ldx() const { return N[0]; }
    virtual bool has_west_column() const { return false; }
};

template <typename T>
struct SolverT : ob_solver {
    using C = typename Cx<T>::type;
    C *S = nullptr, *B = nullptr;
    T *lam[3] = {nullptr, nullptr, nullptr};
    C *tw_f[3] = {nullptr, nullptr, nullptr}, *tw_b[3] = {nullptr, nullptr, nullptr};
    T *diag = nullptr, *lower = nullptr, *tscr = nullptr;
    cufftHandle plan_x = 0, plan_y = 0, plan_z = 0, plan_xy = 0, plan_r2c = 0, plan_c2r = 0;
    bool has_x = false, has_y = false, has_z = false, has_xy = false;
    bool r2c = false;   // fully periodic: 3-D real-to-complex / complex-to-real transforms on the half spectrum
    bool rdct = false;  // x, y Periodic + z Bounded: real DCT along z, real-to-complex 2-D transform in (x, y)
    bool rtri = false;  // x, y Periodic + stretched z: real-to-complex 2-D transform, Thomas sweep on the half spectrum
    C *Cz = nullptr, *tw_zf = nullptr, *tw_zb = nullptr;
    int Nzh = 0;
    cufftHandle plan_zr2c = 0, plan_zc2r = 0, plan_xyr2c = 0, plan_xyc2r = 0;
    T *Rr = nullptr;
    int Nxh = 0;
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
        for (int d = 0; d < 3; d++) { N[d] = g->N[d]; topo[d] = g->topology[d]; L[d] = g->L[d]; }
        tridiag = g->dzf_host != nullptr;
        if (tridiag && topo[2] != OB_BOUNDED) return fail(OB_ERR_UNSUPPORTED, "FourierTridiagonalPoissonSolver needs a Bounded stretched direction");
        const long n = (long)N[0] * N[1] * N[2];
        r2c = !tridiag && topo[0] == OB_PERIODIC && topo[1] == OB_PERIODIC && topo[2] == OB_PERIODIC && N[0] > 1 && N[1] > 1 && N[2] > 1 &&
              !getenv("OB_SOLVER_NO_R2C");
        rdct = !tridiag && topo[0] == OB_PERIODIC && topo[1] == OB_PERIODIC && topo[2] == OB_BOUNDED && N[0] > 1 && N[1] > 1 && N[2] > 1 &&
               !getenv("OB_SOLVER_NO_R2C");
        rtri = tridiag && topo[0] == OB_PERIODIC && topo[1] == OB_PERIODIC && N[0] > 1 && N[1] > 1 && !getenv("OB_SOLVER_NO_R2C");
        if (rtri) {
            // FourierTridiagonalPoissonSolver on the half spectrum of the real rhs (fourier_tridiagonal_poisson_solver.jl:199-260)
            Nxh = N[0] / 2 + 1;
            const long nh = (long)Nxh * N[1] * N[2];
            CUDA_TRY(cudaMalloc(&S, sizeof(C) * nh));
            CUDA_TRY(cudaMalloc(&Rr, sizeof(T) * n));
            CUDA_TRY(cudaMemsetAsync(Rr, 0, sizeof(T) * n, ctx->stream));
            std::vector<T> lx(N[0]), ly(N[1]);
            for (int d = 0; d < 2; d++)
                for (int i = 0; i < N[d]; i++) { double sn = 2 * sin(i * M_PI / N[d]) / (L[d] / N[d]); (d == 0 ? lx : ly)[i] = (T)(sn * sn); }
            const int Nz = N[2], Hz = g->H[2];
            const T *dzf = (const T *)g->dzf_host, *dzc = (const T *)g->dzc_host;
            auto DZF = [&](int k) { return dzf[k + Hz]; };
            auto DZC = [&](int k) { return dzc[k + Hz - 1]; };
            std::vector<T> D((size_t)nh), low(std::max(1, Nz - 1));
            for (int k = 1; k <= Nz; k++)
                for (int j = 0; j < N[1]; j++)
                    for (int i = 0; i < Nxh; i++) {
                        T l = lx[i] + ly[j];
                        T v;
                        if (k == 1) v = (T)-1 / DZF(2) - DZC(1) * l;
                        else if (k == Nz) v = (T)-1 / DZF(Nz) - DZC(Nz) * l;
                        else v = -((T)1 / DZF(k + 1) + (T)1 / DZF(k)) - DZC(k) * l;
                        D[i + (size_t)Nxh * (j + (size_t)N[1] * (k - 1))] = v;
                    }
            for (int q = 1; q <= Nz - 1; q++) low[q - 1] = (T)1 / DZF(q + 1);
            CUDA_TRY(cudaMalloc(&diag, sizeof(T) * nh)); CUDA_TRY(cudaMalloc(&tscr, sizeof(T) * nh));
            CUDA_TRY(cudaMalloc(&lower, sizeof(T) * low.size()));
            CUDA_TRY(cudaMemcpy(diag, D.data(), sizeof(T) * nh, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(lower, low.data(), sizeof(T) * low.size(), cudaMemcpyHostToDevice));
            constexpr cufftType FWD = std::is_same<T, double>::value ? CUFFT_D2Z : CUFFT_R2C;
            constexpr cufftType BWD = std::is_same<T, double>::value ? CUFFT_Z2D : CUFFT_C2R;
            const int plane = N[0] * N[1];
            int nxy[2] = {N[1], N[0]}, exy_r[2] = {N[1], N[0]}, exy_c[2] = {N[1], Nxh};
            CUFFT_TRY(cufftPlanMany(&plan_xyr2c, 2, nxy, exy_r, 1, plane, exy_c, 1, Nxh * N[1], FWD, N[2]));
            CUFFT_TRY(cufftPlanMany(&plan_xyc2r, 2, nxy, exy_c, 1, Nxh * N[1], exy_r, 1, plane, BWD, N[2]));
            CUFFT_TRY(cufftSetStream(plan_xyr2c, ctx->stream)); CUFFT_TRY(cufftSetStream(plan_xyc2r, ctx->stream));
            scale_ = 1.0 / ((double)N[0] * N[1]);
            return OB_OK;
        }
        if (rdct) {
            Nxh = N[0] / 2 + 1; Nzh = N[2] / 2 + 1;
            const long nh = (long)Nxh * N[1] * N[2], nz2 = (long)N[0] * N[1] * Nzh;
            CUDA_TRY(cudaMalloc(&S, sizeof(C) * nh));
            CUDA_TRY(cudaMalloc(&Cz, sizeof(C) * nz2));
            CUDA_TRY(cudaMalloc(&Rr, sizeof(T) * n));
            CUDA_TRY(cudaMemsetAsync(Rr, 0, sizeof(T) * n, ctx->stream));
            for (int d = 0; d < 3; d++) {
                std::vector<T> h(N[d]);
                for (int i = 0; i < N[d]; i++) {
                    double sn = 2 * sin(i * M_PI / ((d == 2 ? 2.0 : 1.0) * N[d])) / (L[d] / N[d]);
                    h[i] = (T)(sn * sn);
                }
                CUDA_TRY(cudaMalloc(&lam[d], sizeof(T) * N[d]));
                CUDA_TRY(cudaMemcpy(lam[d], h.data(), sizeof(T) * N[d], cudaMemcpyHostToDevice));
            }
            std::vector<C> f(N[2]), b(Nzh);
            for (int k = 0; k < N[2]; k++) { double a = -M_PI * k / (2.0 * N[2]); f[k].x = (T)cos(a); f[k].y = (T)sin(a); }
            for (int k = 0; k < Nzh; k++) { double a = M_PI * k / (2.0 * N[2]); b[k].x = (T)(0.5 * cos(a)); b[k].y = (T)(0.5 * sin(a)); }
            CUDA_TRY(cudaMalloc(&tw_zf, sizeof(C) * N[2])); CUDA_TRY(cudaMalloc(&tw_zb, sizeof(C) * Nzh));
            CUDA_TRY(cudaMemcpy(tw_zf, f.data(), sizeof(C) * N[2], cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(tw_zb, b.data(), sizeof(C) * Nzh, cudaMemcpyHostToDevice));
            constexpr cufftType FWD = std::is_same<T, double>::value ? CUFFT_D2Z : CUFFT_R2C;
            constexpr cufftType BWD = std::is_same<T, double>::value ? CUFFT_Z2D : CUFFT_C2R;
            const int plane = N[0] * N[1];
            int nzz[1] = {N[2]}, nre[1] = {N[2]}, nco[1] = {Nzh};
            CUFFT_TRY(cufftPlanMany(&plan_zr2c, 1, nzz, nre, plane, 1, nco, plane, 1, FWD, plane));
            CUFFT_TRY(cufftPlanMany(&plan_zc2r, 1, nzz, nco, plane, 1, nre, plane, 1, BWD, plane));
            int nxy[2] = {N[1], N[0]}, exy_r[2] = {N[1], N[0]}, exy_c[2] = {N[1], Nxh};
            CUFFT_TRY(cufftPlanMany(&plan_xyr2c, 2, nxy, exy_r, 1, plane, exy_c, 1, Nxh * N[1], FWD, N[2]));
            CUFFT_TRY(cufftPlanMany(&plan_xyc2r, 2, nxy, exy_c, 1, Nxh * N[1], exy_r, 1, plane, BWD, N[2]));
            for (cufftHandle h : {plan_zr2c, plan_zc2r, plan_xyr2c, plan_xyc2r}) CUFFT_TRY(cufftSetStream(h, ctx->stream));
            scale_ = 1.0 / ((double)N[0] * N[1] * N[2]);
            return OB_OK;
        }
        if (r2c) {
            // The rhs is real, so a real-to-complex 3-D transform carries half the bytes of the reference's complex
            // in-place FFTs (fft_based_poisson_solver.jl:94-124); same eigenvalue division on the half spectrum i <= Nx/2.
            Nxh = N[0] / 2 + 1;
            const long nh = (long)Nxh * N[1] * N[2];
            CUDA_TRY(cudaMalloc(&S, sizeof(C) * nh));
            CUDA_TRY(cudaMalloc(&Rr, sizeof(T) * n));
            CUDA_TRY(cudaMemsetAsync(Rr, 0, sizeof(T) * n, ctx->stream));
            for (int d = 0; d < 3; d++) {
                std::vector<T> h(N[d]);
                for (int i = 0; i < N[d]; i++) { double sn = 2 * sin(i * M_PI / N[d]) / (L[d] / N[d]); h[i] = (T)(sn * sn); }
                CUDA_TRY(cudaMalloc(&lam[d], sizeof(T) * N[d]));
                CUDA_TRY(cudaMemcpy(lam[d], h.data(), sizeof(T) * N[d], cudaMemcpyHostToDevice));
            }
            constexpr cufftType FWD = std::is_same<T, double>::value ? CUFFT_D2Z : CUFFT_R2C;
            constexpr cufftType BWD = std::is_same<T, double>::value ? CUFFT_Z2D : CUFFT_C2R;
            CUFFT_TRY(cufftPlan3d(&plan_r2c, N[2], N[1], N[0], FWD));
            CUFFT_TRY(cufftPlan3d(&plan_c2r, N[2], N[1], N[0], BWD));
            CUFFT_TRY(cufftSetStream(plan_r2c, ctx->stream));
            CUFFT_TRY(cufftSetStream(plan_c2r, ctx->stream));
            scale_ = 1.0 / ((double)N[0] * N[1] * N[2]);
            return OB_OK;
        }
        CUDA_TRY(cudaMalloc(&S, sizeof(C) * n));
        CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(C) * n, ctx->stream));
        bool any_bounded = false;
        const int ndim_t = tridiag ? 2 : 3;
        for (int d = 0; d < ndim_t; d++) any_bounded |= topo[d] == OB_BOUNDED;
        // work array of the DCT passes, and of a Periodic y transform that is not part of a 2-D (x, y) plan
        const bool y_alone = topo[1] == OB_PERIODIC && N[1] > 1 && !(topo[0] == OB_PERIODIC && N[0] > 1);
        if (any_bounded || y_alone) CUDA_TRY(cudaMalloc(&B, sizeof(C) * n));
        // eigenvalues (poisson_eigenvalues.jl:8-32), computed in Float64 then converted to FT
        for (int d = 0; d < 3; d++) {
            std::vector<T> h(N[d]);
            for (int i = 0; i < N[d]; i++) {
                double v = 0;
                if (topo[d] == OB_PERIODIC) { double s = 2 * sin(i * M_PI / N[d]) / (L[d] / N[d]); v = s * s; }
                else if (topo[d] == OB_BOUNDED) { double s = 2 * sin(i * M_PI / (2.0 * N[d])) / (L[d] / N[d]); v = s * s; }
                h[i] = (T)v;
            }
            CUDA_TRY(cudaMalloc(&lam[d], sizeof(T) * N[d]));
            CUDA_TRY(cudaMemcpy(lam[d], h.data(), sizeof(T) * N[d], cudaMemcpyHostToDevice));
            if (topo[d] == OB_BOUNDED && d < ndim_t) {  // twiddles ω_4N^{±k} (discrete_transforms.jl:48-78)
                std::vector<C> f(N[d]), b(N[d]);
                for (int k = 0; k < N[d]; k++) {
                    double a = -2 * M_PI * k / (4.0 * N[d]);
                    f[k].x = (T)cos(a); f[k].y = (T)sin(a);
                    b[k].x = (T)cos(-a); b[k].y = (T)sin(-a);
                }
                b[0].x *= (T)0.5; b[0].y *= (T)0.5;
                CUDA_TRY(cudaMalloc(&tw_f[d], sizeof(C) * N[d]));
                CUDA_TRY(cudaMalloc(&tw_b[d], sizeof(C) * N[d]));
                CUDA_TRY(cudaMemcpy(tw_f[d], f.data(), sizeof(C) * N[d], cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(tw_b[d], b.data(), sizeof(C) * N[d], cudaMemcpyHostToDevice));
            }
        }
        // plans
        auto transformed = [&](int d) { return d < ndim_t && topo[d] != OB_FLAT && N[d] > 1; };
        scale_ = 1.0;
        for (int d = 0; d < 3; d++) if (transformed(d)) scale_ /= N[d];
        if (transformed(0) && transformed(1) && topo[0] == OB_PERIODIC && topo[1] == OB_PERIODIC) {
            int nn[2] = {N[1], N[0]};
            CUFFT_TRY(cufftPlanMany(&plan_xy, 2, nn, nullptr, 1, N[0] * N[1], nullptr, 1, N[0] * N[1], CT, N[2]));
            CUFFT_TRY(cufftSetStream(plan_xy, ctx->stream));
            has_xy = true;
        } else {
            if (transformed(0)) {
                int nn[1] = {N[0]};
                CUFFT_TRY(cufftPlanMany(&plan_x, 1, nn, nn, 1, N[0], nn, 1, N[0], CT, N[1] * N[2]));
                CUFFT_TRY(cufftSetStream(plan_x, ctx->stream));
                has_x = true;
            }
            if (transformed(1)) {  // contiguous y lines of the y-fastest work layout (poisson.cuh: bidx), all Nx Nz of them in one call
                int nn[1] = {N[1]};
                CUFFT_TRY(cufftPlanMany(&plan_y, 1, nn, nn, 1, N[1], nn, 1, N[1], CT, N[0] * N[2]));
                CUFFT_TRY(cufftSetStream(plan_y, ctx->stream));
                has_y = true;
            }
        }
        if (transformed(2)) {
            int nn[1] = {N[2]};
            CUFFT_TRY(cufftPlanMany(&plan_z, 1, nn, nn, N[0] * N[1], 1, nn, N[0] * N[1], 1, CT, N[0] * N[1]));
            CUFFT_TRY(cufftSetStream(plan_z, ctx->stream));
            has_z = true;
        }
        if (tridiag) {
            // main diagonal & lower diagonal (fourier_tridiagonal_poisson_solver.jl:199-229)
            const int Nz = N[2], Hz = g->H[2];
            const T *dzf = (const T *)g->dzf_host, *dzc = (const T *)g->dzc_host;
            auto DZF = [&](int k) { return dzf[k + Hz]; };
            auto DZC = [&](int k) { return dzc[k + Hz - 1]; };
            std::vector<T> lx(N[0]), ly(N[1]);
            CUDA_TRY(cudaMemcpy(lx.data(), lam[0], sizeof(T) * N[0], cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(ly.data(), lam[1], sizeof(T) * N[1], cudaMemcpyDeviceToHost));
            std::vector<T> D((size_t)n), low(std::max(1, Nz - 1));
            for (int k = 1; k <= Nz; k++)
                for (int j = 0; j < N[1]; j++)
                    for (int i = 0; i < N[0]; i++) {
                        T l = lx[i] + ly[j];
                        T v;
                        if (k == 1) v = (T)-1 / DZF(2) - DZC(1) * l;
                        else if (k == Nz) v = (T)-1 / DZF(Nz) - DZC(Nz) * l;
                        else v = -((T)1 / DZF(k + 1) + (T)1 / DZF(k)) - DZC(k) * l;
                        D[i + (size_t)N[0] * (j + (size_t)N[1] * (k - 1))] = v;
                    }
            for (int q = 1; q <= Nz - 1; q++) low[q - 1] = (T)1 / DZF(q + 1);
            CUDA_TRY(cudaMalloc(&diag, sizeof(T) * n));
            CUDA_TRY(cudaMalloc(&tscr, sizeof(T) * n));
            CUDA_TRY(cudaMalloc(&lower, sizeof(T) * low.size()));
            CUDA_TRY(cudaMemcpy(diag, D.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(lower, low.data(), sizeof(T) * low.size(), cudaMemcpyHostToDevice));
        }
        return OB_OK;
    }
    ~SolverT() override {
        cudaFree(S); cudaFree(B);
        for (int d = 0; d < 3; d++) { cudaFree(lam[d]); cudaFree(tw_f[d]); cudaFree(tw_b[d]); }
        cudaFree(diag); cudaFree(lower); cudaFree(tscr);
        if (has_x) cufftDestroy(plan_x);
        if (has_y) cufftDestroy(plan_y);
        if (has_z) cufftDestroy(plan_z);
        if (has_xy) cufftDestroy(plan_xy);
        if (r2c) { cufftDestroy(plan_r2c); cufftDestroy(plan_c2r); cudaFree(Rr); }
        if (rtri) { cufftDestroy(plan_xyr2c); cufftDestroy(plan_xyc2r); cudaFree(Rr); }
        if (rdct) { cufftDestroy(plan_zr2c); cufftDestroy(plan_zc2r); cufftDestroy(plan_xyr2c); cufftDestroy(plan_xyc2r); cudaFree(Rr); cudaFree(Cz); cudaFree(tw_zf); cudaFree(tw_zb); }
    }
    void *storage() override { return (r2c || rdct || rtri) ? (void *)Rr : (void *)S; }
    double scale() override { return scale_; }
    bool real_storage() const override { return r2c || rdct || rtri; }
    bool z_permuted() const override { return rdct; }

    int32_t fft_dim(C *data, int d, int dir) {
        if (d == 0) return exec(plan_x, data, dir);
        if (d == 2) return exec(plan_z, data, dir);
        return exec(plan_y, data, dir);   // `data` is in the y-fastest layout
    }
    // 1-D FFT along a Periodic y outside the 2-D (x, y) plan: to the y-fastest layout, one batched call, and back
    int32_t fft_y_periodic(int dir) {
        const long n = (long)N[0] * N[1] * N[2];
        transpose_y_kernel<C><<<nblk(n, 256), 256, 0, ctx->stream>>>(S, B, N[0], N[1], N[2], 0);
        OB_TRY(exec(plan_y, B, dir));
        transpose_y_kernel<C><<<nblk(n, 256), 256, 0, ctx->stream>>>(B, S, N[0], N[1], N[2], 1);
        launches += 2;
        return OB_OK;
    }
    int32_t solve_in_storage() override {
        const long n = (long)N[0] * N[1] * N[2];
        const unsigned nb = nblk(n, 256);
        cudaStream_t st = ctx->stream;
        if (rtri) {
            constexpr bool DBL = std::is_same<T, double>::value;
            if constexpr (DBL) CUFFT_TRY(cufftExecD2Z(plan_xyr2c, Rr, S)); else CUFFT_TRY(cufftExecR2C(plan_xyr2c, (cufftReal *)Rr, (cufftComplex *)S));
            dim3 grid(nblk(Nxh, 128), N[1]);
            thomas_kernel<T, C><<<grid, 128, 0, st>>>(S, lower, lower, diag, tscr, Nxh, N[1], N[2], (T)(10 * std::numeric_limits<T>::epsilon()), 1);
            if constexpr (DBL) CUFFT_TRY(cufftExecZ2D(plan_xyc2r, S, Rr)); else CUFFT_TRY(cufftExecC2R(plan_xyc2r, (cufftComplex *)S, (cufftReal *)Rr));
            launches += 3;
            CUDA_TRY(cudaGetLastError());
            return OB_OK;
        }
        if (rdct) {
            const long nz2 = (long)N[0] * N[1] * Nzh, nh = (long)Nxh * N[1] * N[2];
            constexpr bool DBL = std::is_same<T, double>::value;
            if constexpr (DBL) CUFFT_TRY(cufftExecD2Z(plan_zr2c, Rr, Cz)); else CUFFT_TRY(cufftExecR2C(plan_zr2c, (cufftReal *)Rr, (cufftComplex *)Cz));
            dct_z_real_fwd_kernel<T, C><<<nb, 256, 0, st>>>(Cz, Rr, tw_zf, N[0], N[1], N[2], Nzh);
            if constexpr (DBL) CUFFT_TRY(cufftExecD2Z(plan_xyr2c, Rr, S)); else CUFFT_TRY(cufftExecR2C(plan_xyr2c, (cufftReal *)Rr, (cufftComplex *)S));
            eigen_divide_kernel<T, C><<<nblk(nh, 256), 256, 0, st>>>(S, lam[0], lam[1], lam[2], Nxh, N[1], N[2]);
            if constexpr (DBL) CUFFT_TRY(cufftExecZ2D(plan_xyc2r, S, Rr)); else CUFFT_TRY(cufftExecC2R(plan_xyc2r, (cufftComplex *)S, (cufftReal *)Rr));
            dct_z_real_bwd_kernel<T, C><<<nblk(nz2, 256), 256, 0, st>>>(Rr, Cz, tw_zb, N[0], N[1], N[2], Nzh);
            if constexpr (DBL) CUFFT_TRY(cufftExecZ2D(plan_zc2r, Cz, Rr)); else CUFFT_TRY(cufftExecC2R(plan_zc2r, (cufftComplex *)Cz, (cufftReal *)Rr));
            launches += 7;
            CUDA_TRY(cudaGetLastError());
            return OB_OK;
        }
        if (r2c) {
            if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecD2Z(plan_r2c, Rr, S));
            else CUFFT_TRY(cufftExecR2C(plan_r2c, Rr, S));
            const long nh = (long)Nxh * N[1] * N[2];
            eigen_divide_kernel<T, C><<<nblk(nh, 256), 256, 0, st>>>(S, lam[0], lam[1], lam[2], Nxh, N[1], N[2]);
            if constexpr (std::is_same<T, double>::value) CUFFT_TRY(cufftExecZ2D(plan_c2r, S, Rr));
            else CUFFT_TRY(cufftExecC2R(plan_c2r, S, Rr));
            launches += 3;
            CUDA_TRY(cudaGetLastError());
            return OB_OK;
        }
        const int ndim_t = tridiag ? 2 : 3;
        auto transformed = [&](int d) { return d < ndim_t && topo[d] != OB_FLAT && N[d] > 1; };
        // forward: Bounded dims first (plan_transforms.jl:160-199), then Periodic
        for (int d = 0; d < ndim_t; d++)
            if (transformed(d) && topo[d] == OB_BOUNDED) {
                permute_kernel<C><<<nb, 256, 0, st>>>(S, B, N[0], N[1], N[2], d, d == 1);
                OB_TRY(fft_dim(B, d, CUFFT_FORWARD));
                twiddle_fwd_kernel<T, C><<<nb, 256, 0, st>>>(B, S, tw_f[d], N[0], N[1], N[2], d, d == 1);
                launches += 2;
            }
        if (has_xy) OB_TRY(exec(plan_xy, S, CUFFT_FORWARD));
        else {
            if (transformed(0) && topo[0] == OB_PERIODIC) OB_TRY(fft_dim(S, 0, CUFFT_FORWARD));
            if (transformed(1) && topo[1] == OB_PERIODIC) OB_TRY(fft_y_periodic(CUFFT_FORWARD));
        }
        if (transformed(2) && topo[2] == OB_PERIODIC) OB_TRY(fft_dim(S, 2, CUFFT_FORWARD));
        if (tridiag) {
            dim3 grid(nblk(N[0], 128), N[1]);
            thomas_kernel<T, C><<<grid, 128, 0, st>>>(S, lower, lower, diag, tscr, N[0], N[1], N[2], (T)(10 * std::numeric_limits<T>::epsilon()), 1);
        } else {
            eigen_divide_kernel<T, C><<<nb, 256, 0, st>>>(S, lam[0], lam[1], lam[2], N[0], N[1], N[2]);
        }
        launches++;
        // backward: Periodic first, then Bounded
        if (transformed(2) && topo[2] == OB_PERIODIC) OB_TRY(fft_dim(S, 2, CUFFT_INVERSE));
        if (has_xy) OB_TRY(exec(plan_xy, S, CUFFT_INVERSE));
        else {
            if (transformed(1) && topo[1] == OB_PERIODIC) OB_TRY(fft_y_periodic(CUFFT_INVERSE));
            if (transformed(0) && topo[0] == OB_PERIODIC) OB_TRY(fft_dim(S, 0, CUFFT_INVERSE));
        }
        for (int d = ndim_t - 1; d >= 0; d--)
            if (transformed(d) && topo[d] == OB_BOUNDED) {
                twiddle_bwd_kernel<T, C><<<nb, 256, 0, st>>>(S, B, tw_b[d], N[0], N[1], N[2], d, d == 1);
                OB_TRY(fft_dim(B, d, CUFFT_INVERSE));
                unpermute_kernel<C><<<nb, 256, 0, st>>>(B, S, N[0], N[1], N[2], d, d == 1);
                launches += 2;
            }
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
};

#include "dist_solver.cuh"

template <typename T>
static int32_t make_solver(ob_ctx *ctx, const ob_grid_desc *grid, ob_solver **out) {
    int32_t st;
    ob_solver *s;
    if (ctx->world > 1) { auto *p = new DistSolverT<T>(); st = p->init(ctx, grid); s = p; }
    else { auto *p = new SolverT<T>(); st = p->init(ctx, grid); s = p; }
    if (st != OB_OK) { delete s; return st; }
    *out = s;
    return OB_OK;
}

extern "C" int32_t ob_solver_create(ob_ctx *ctx, const ob_grid_desc *grid, ob_solver **out) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    ob_solver *s = nullptr;
    int32_t st;
    if (grid->float_type == OB_F64) st = make_solver<double>(ctx, grid, &s);
    else st = make_solver<float>(ctx, grid, &s);
    if (st != OB_OK) return st;
    *out = s;
    return OB_OK;
}
extern "C" int32_t ob_solver_destroy(ob_solver *s) { delete s; return OB_OK; }

template <typename T, typename C>
__global__ void pack_real_kernel(const T *r, C *S, long n) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) { C v; v.x = r[t]; v.y = 0; S[t] = v; }
}
template <typename T, typename C>
__global__ void unpack_real_kernel(const C *S, T *r, long n, T scale) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) r[t] = S[t].x * scale;
}
// copy with an optional Makhoul permutation of the levels: forward = 1 writes level k at its permuted position,
// forward = 0 reads it from there (Nz = 0: plain copy); out = in * scale
template <typename T>
__global__ void zperm_copy_kernel(const T *in, T *out, long n, long plane, int Nz, int forward, T scale) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long o = t;
    if (Nz > 0) {
        const long k = t / plane, ij = t - k * plane;
        const long kp = (k & 1) ? (long)Nz - 1 - (k >> 1) : (k >> 1);
        o = ij + kp * plane;
    }
    if (forward) out[o] = in[t] * scale;
    else out[t] = in[o] * scale;
}
extern "C" int32_t ob_poisson_solve(ob_solver *s, const void *rhs, void *phi) {
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    const long n = (long)s->N[0] * s->N[1] * s->N[2];
    cudaStream_t st = s->ctx->stream;
    if (s->real_storage()) {
        const long plane = (long)s->N[0] * s->N[1];
        const int zp = s->z_permuted() ? s->N[2] : 0;
        if (s->ft == OB_F64) zperm_copy_kernel<double><<<nblk(n, 256), 256, 0, st>>>((const double *)rhs, (double *)s->storage(), n, plane, zp, 1, 1.0);
        else zperm_copy_kernel<float><<<nblk(n, 256), 256, 0, st>>>((const float *)rhs, (float *)s->storage(), n, plane, zp, 1, 1.0f);
        OB_TRY(s->solve_in_storage());
        if (s->ft == OB_F64) zperm_copy_kernel<double><<<nblk(n, 256), 256, 0, st>>>((const double *)s->storage(), (double *)phi, n, plane, zp, 0, s->scale());
        else zperm_copy_kernel<float><<<nblk(n, 256), 256, 0, st>>>((const float *)s->storage(), (float *)phi, n, plane, zp, 0, (float)s->scale());
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    if (s->ft == OB_F64) pack_real_kernel<double, double2><<<nblk(n, 256), 256, 0, st>>>((const double *)rhs, (double2 *)s->storage(), n);
    else pack_real_kernel<float, float2><<<nblk(n, 256), 256, 0, st>>>((const float *)rhs, (float2 *)s->storage(), n);
    OB_TRY(s->solve_in_storage());
    if (s->ft == OB_F64) unpack_real_kernel<double, double2><<<nblk(n, 256), 256, 0, st>>>((const double2 *)s->storage(), (double *)phi, n, s->scale());
    else unpack_real_kernel<float, float2><<<nblk(n, 256), 256, 0, st>>>((const float2 *)s->storage(), (float *)phi, n, (float)s->scale());
    CUDA_TRY(cudaGetLastError());
    return OB_OK;
}

extern "C" int32_t ob_batched_tridiagonal_solve(ob_ctx *ctx, int32_t ft, int32_t is_complex, int32_t Nx, int32_t Ny, int32_t Nz,
                                                const void *a, const void *b, const void *c, const void *f, void *phi, void *scratch) {
    if (!is_complex) return fail(OB_ERR_UNSUPPORTED, "ob_batched_tridiagonal_solve: real right-hand sides: pass complex with zero imaginary part");
    const long n = (long)Nx * Ny * Nz;
    dim3 grid(nblk(Nx, 128), Ny);
    if (ft == OB_F64) {
        if (phi != f) CUDA_TRY(cudaMemcpyAsync(phi, f, sizeof(double2) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        thomas_kernel<double, double2><<<grid, 128, 0, ctx->stream>>>((double2 *)phi, (const double *)a, (const double *)c, (const double *)b, (double *)scratch, Nx, Ny, Nz,
                                                                      10 * std::numeric_limits<double>::epsilon(), 0);
    } else {
        if (phi != f) CUDA_TRY(cudaMemcpyAsync(phi, f, sizeof(float2) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        thomas_kernel<float, float2><<<grid, 128, 0, ctx->stream>>>((float2 *)phi, (const float *)a, (const float *)c, (const float *)b, (float *)scratch, Nx, Ny, Nz,
                                                                    10 * std::numeric_limits<float>::epsilon(), 0);
    }
    CUDA_TRY(cudaGetLastError());
    return OB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------------------
enum Phase { PH_HALO = 0, PH_CLOSURE, PH_HYDRO, PH_TENDENCY, PH_UPDATE, PH_SOURCE, PH_SOLVE, PH_CORRECT, PH_COUNT };
static const char *PHASE_NAMES[PH_COUNT] = {"halo", "closure_fields", "hydrostatic_pressure", "tendencies", "update", "poisson_source", "poisson_solve", "pressure_correct"};

struct FieldInfo {
    void *ptr = nullptr;
    int loc[3] = {0, 0, 0};  // 1 = Face
    int P[3] = {1, 1, 1}, n[3] = {1, 1, 1}, o[3] = {0, 0, 0};
    ob_bc_desc bc;
    const void *bc_array[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // ob_model_set_bc_array
    bool exists = false;
    long count() const { return (long)P[0] * P[1] * P[2]; }
};

struct ob_model {
    ob_ctx *ctx = nullptr;
    ob_model_desc desc;
    int64_t launches = 0;
    bool timing = false;
    int opt_tendency_kernel = 0;  // OB_OPT_TENDENCY_KERNEL: 0 auto, 1 generic (one thread per cell), 2 marching
    int opt_overlap = 0;          // OB_OPT_OVERLAP_HALO: distributed update_state computes interior tendency tiles while x halos are in flight
                                  // (off by default: at 256^3 per GPU the split launch costs 0.2 ms/step more than the exchange it hides)
    int opt_fuse = 1;             // OB_OPT_FUSE_PROJECTION: 1 fused single-device projection, 0 the reference kernel sequence
    int opt_vector = 1;           // OB_OPT_VECTOR_STREAMS: 1 128-bit forms of update / source / projection kernels (streaming.cuh), 0 one cell per thread
    double phase_ms[PH_COUNT] = {0};
    int64_t phase_calls[PH_COUNT] = {0};
    struct Ev { int phase; cudaEvent_t a, b; };
    std::vector<Ev> pending;
    std::vector<cudaEvent_t> pool;
    virtual ~ob_model() {}
    virtual int32_t bind(int32_t id, void *p) = 0;
    virtual int32_t fill_halo(int32_t id, int32_t fill_normal) = 0;
    virtual int32_t update_state() = 0;
    virtual int32_t compute_tendencies() = 0;
    virtual int32_t compute_closure_fields() = 0;
    virtual int32_t update_hydrostatic_pressure() = 0;
    virtual int32_t set_bc_array(int id, int side, const void *p) = 0;
    virtual int32_t rk3_substep(double dt, double gamma, double zeta, int has_zeta, bool cache) = 0;
    virtual int32_t ab2_step(double dt, double chi, bool cache) = 0;
    virtual int32_t cache_tendencies() = 0;
    virtual int32_t compute_pressure_correction(double dtau) = 0;
    virtual int32_t make_pressure_correction(double dtau) = 0;
    virtual int32_t time_step_rk3(double dt, int first) = 0;
    virtual int32_t time_step_ab2(double dt, int euler, int first) = 0;
    virtual int32_t advection_timescale(double *tau) = 0;

    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int ph, cudaEvent_t &a) { if (timing) { a = get_event(); cudaEventRecord(a, ctx->stream); } (void)ph; }
    void end(int ph, cudaEvent_t a) {
        if (timing) { cudaEvent_t b = get_event(); cudaEventRecord(b, ctx->stream); pending.push_back({ph, a, b}); }
    }
    void collect() {
        if (pending.empty()) return;
        cudaStreamSynchronize(ctx->stream);
        for (auto &e : pending) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e.a, e.b);
            phase_ms[e.phase] += ms;
            phase_calls[e.phase]++;
            pool.push_back(e.a);
            pool.push_back(e.b);
        }
        pending.clear();
    }
};
struct PhaseScope {
    ob_model *m; int ph; cudaEvent_t a = nullptr;
    PhaseScope(ob_model *m_, int ph_) : m(m_), ph(ph_) { m->begin(ph, a); }
    ~PhaseScope() { m->end(ph, a); }
};

template <typename T>
struct ModelT : ob_model {
    GridD<T> g;
    T *d_dzf = nullptr, *d_dzc = nullptr, *d_rdzf = nullptr, *d_rdzc = nullptr, *d_rvf = nullptr, *d_rvc = nullptr;
    std::vector<T> h_dzf, h_dzc;
    int Hz_ = 0;
    FieldInfo F[128];
    ob_solver *solver = nullptr;
    int ntr = 0, ncl = 0;
    double *d_partial = nullptr;
    bool dist = false;
    T *d_halo_buf = nullptr;
    size_t halo_buf_elems = 0;
    // peer-to-peer halo exchange state (CUDA IPC): staging [parity][side(w,e)][cap] + flags [parity][side]
    bool p2p_halo = false;
    T *d_stage = nullptr; int *d_flags = nullptr; unsigned *d_blockctr = nullptr;
    T *west_stage = nullptr, *east_stage = nullptr; int *west_flags = nullptr, *east_flags = nullptr;
    size_t stage_cap = 0;
    int halo_epoch = 0;

    ~ModelT() override {
        cudaFree(d_dzf); cudaFree(d_dzc); cudaFree(d_partial);
        cudaFree(d_rdzf); cudaFree(d_rdzc); cudaFree(d_rvf); cudaFree(d_rvc); cudaFree(d_halo_buf);
        if (p2p_halo) {
            cudaStreamSynchronize(ctx->stream);
            const int R = ctx->world, west = (ctx->rank + R - 1) % R, east = (ctx->rank + 1) % R;
            cudaIpcCloseMemHandle(west_stage); cudaIpcCloseMemHandle(west_flags);
            if (east != west) { cudaIpcCloseMemHandle(east_stage); cudaIpcCloseMemHandle(east_flags); }
        }
        cudaFree(d_stage); cudaFree(d_flags); cudaFree(d_blockctr); cudaFree(d_ivd); cudaFree(d_amd_tab);
        for (int m = 0; m < OB_MAXCL; m++) {
            T *dq[8] = {dynw[m].ub, dynw[m].vb, dynw[m].wb, dynw[m].Sg, dynw[m].Sb, dynw[m].LM, dynw[m].MM, dynw[m].J};
            for (int q = 0; q < 8; q++) cudaFree(dq[q]);
        }
        delete solver;
        for (auto e : pool) cudaEventDestroy(e);
    }

    // ---- geometry -------------------------------------------------------------------------------------------
    void setup_field(int id, int lx, int ly, int lz, const ob_bc_desc &bc) {
        FieldInfo &f = F[id];
        f.exists = true;
        f.loc[0] = lx; f.loc[1] = ly; f.loc[2] = lz;
        for (int d = 0; d < 3; d++) {
            f.n[d] = g.N[d] + ((f.loc[d] && g.topo[d] == BOUNDED) ? 1 : 0);
            f.P[d] = f.n[d] + 2 * g.H[d];
            f.o[d] = g.H[d];
        }
        f.bc = bc;
    }
    // G-role swap (time_step_rk3): while set, the tendency fields Gⁿ and G⁻ trade places, so that "G⁻ <- Gⁿ" between two
    // stages is a change of roles instead of a copy
    bool gswap = false;
    Fld<T> fld(int id) const {
        if (gswap) {
            if (id >= OB_FIELD_GN0 && id < OB_FIELD_GN0 + 3 + OB_MAX_TRACERS) id += OB_FIELD_GM0 - OB_FIELD_GN0;
            else if (id >= OB_FIELD_GM0 && id < OB_FIELD_GM0 + 3 + OB_MAX_TRACERS) id -= OB_FIELD_GM0 - OB_FIELD_GN0;
        }
        const FieldInfo &f = F[id];
        Fld<T> v;
        v.p = (T *)f.ptr;
        v.sy = f.P[0];
        v.sz = (long)f.P[0] * f.P[1];
        v.off = (long)(f.o[0] - 1) + (long)(f.o[1] - 1) * v.sy + (long)(f.o[2] - 1) * v.sz;
        return v;
    }
    T hDZF(int k) const { return h_dzf.empty() ? g.dz : h_dzf[k + Hz_]; }
    T hDZC(int k) const { return h_dzc.empty() ? g.dz : h_dzc[k + Hz_ - 1]; }

    int32_t init(ob_ctx *c, const ob_model_desc *d) {
        ctx = c;
        desc = *d;
        const ob_grid_desc &gd = d->grid;
        for (int k = 0; k < 3; k++) {
            g.N[k] = gd.N[k]; g.topo[k] = gd.topology[k];
            g.H[k] = gd.topology[k] == OB_FLAT ? 0 : gd.H[k];
            if (gd.topology[k] == OB_FLAT && gd.N[k] != 1) return fail(OB_ERR_INVALID, "Flat dimension %d must have size 1", k);
        }
        g.dx = g.topo[0] == FLAT ? (T)1 : (T)gd.d[0];
        g.dy = g.topo[1] == FLAT ? (T)1 : (T)gd.d[1];
        g.dz = g.topo[2] == FLAT ? (T)1 : (T)gd.d[2];
        g.dzf = g.dzc = nullptr;
        g.rdzf = g.rdzc = g.rvf = g.rvc = nullptr;
        Hz_ = g.H[2];
        if (gd.dzf_host) {
            if (!gd.dzc_host) return fail(OB_ERR_INVALID, "dzf given without dzc");
            h_dzf.assign((const T *)gd.dzf_host, (const T *)gd.dzf_host + gd.n_dzf);
            h_dzc.assign((const T *)gd.dzc_host, (const T *)gd.dzc_host + gd.n_dzc);
            CUDA_TRY(cudaMalloc(&d_dzf, sizeof(T) * gd.n_dzf));
            CUDA_TRY(cudaMalloc(&d_dzc, sizeof(T) * gd.n_dzc));
            CUDA_TRY(cudaMemcpy(d_dzf, h_dzf.data(), sizeof(T) * gd.n_dzf, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_dzc, h_dzc.data(), sizeof(T) * gd.n_dzc, cudaMemcpyHostToDevice));
            g.dzf = d_dzf + Hz_;       // logical k -> dzf[k + Hz]
            g.dzc = d_dzc + Hz_ - 1;   // logical k -> dzc[k + Hz - 1]
            // reciprocal metrics, same IEEE divisions as the reference evaluates per call
            std::vector<T> rf(gd.n_dzf), rc(gd.n_dzc), vf(gd.n_dzf), vc(gd.n_dzc);
            for (int q = 0; q < gd.n_dzf; q++) { rf[q] = (T)1 / h_dzf[q]; vf[q] = (T)1 / ((g.dx * g.dy) * h_dzf[q]); }
            for (int q = 0; q < gd.n_dzc; q++) { rc[q] = (T)1 / h_dzc[q]; vc[q] = (T)1 / ((g.dx * g.dy) * h_dzc[q]); }
            CUDA_TRY(cudaMalloc(&d_rdzf, sizeof(T) * gd.n_dzf)); CUDA_TRY(cudaMalloc(&d_rvf, sizeof(T) * gd.n_dzf));
            CUDA_TRY(cudaMalloc(&d_rdzc, sizeof(T) * gd.n_dzc)); CUDA_TRY(cudaMalloc(&d_rvc, sizeof(T) * gd.n_dzc));
            CUDA_TRY(cudaMemcpy(d_rdzf, rf.data(), sizeof(T) * gd.n_dzf, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_rvf, vf.data(), sizeof(T) * gd.n_dzf, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_rdzc, rc.data(), sizeof(T) * gd.n_dzc, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(d_rvc, vc.data(), sizeof(T) * gd.n_dzc, cudaMemcpyHostToDevice));
            g.rdzf = d_rdzf + Hz_; g.rvf = d_rvf + Hz_;
            g.rdzc = d_rdzc + Hz_ - 1; g.rvc = d_rvc + Hz_ - 1;
        }
        g.rdx = (T)1 / g.dx; g.rdy = (T)1 / g.dy; g.rdz = (T)1 / g.dz;
        g.rvol = (T)1 / ((g.dx * g.dy) * g.dz);
        ntr = d->n_tracers;
        ncl = d->n_closures;
        if (ntr > OB_MAXTR || ncl > OB_MAXCL) return fail(OB_ERR_INVALID, "too many tracers/closures");
        // scope validation: anything else must be rejected, never silently approximated
        // every order the reference builds buffers for (src/Advection/Advection.jl:52: buffers 1 .. 6)
        if (d->advection_kind == OB_ADV_WENO && (d->advection_order < 3 || d->advection_order > 11 || d->advection_order % 2 == 0))
            return fail(OB_ERR_UNSUPPORTED, "WENO order %d not supported (3, 5, 7, 9, 11)", d->advection_order);
        if (d->advection_kind == OB_ADV_CENTERED && (d->advection_order < 2 || d->advection_order > 12 || d->advection_order % 2))
            return fail(OB_ERR_UNSUPPORTED, "Centered order %d not supported (2, 4, ..., 12)", d->advection_order);
        const int nb = d->advection_kind == OB_ADV_WENO ? (d->advection_order + 1) / 2 : d->advection_kind == OB_ADV_CENTERED ? d->advection_order / 2 : 0;
        for (int k = 0; k < 3; k++) {
            if (g.topo[k] != FLAT && g.N[k] < nb) return fail(OB_ERR_UNSUPPORTED, "grid size %d along %d is smaller than the advection buffer %d (adapt_advection_order path)", g.N[k], k, nb);
            if (g.topo[k] != FLAT && g.H[k] < nb) return fail(OB_ERR_INVALID, "halo %d along %d is smaller than the advection buffer %d", g.H[k], k, nb);
            if (g.topo[k] != FLAT && g.H[k] < 1) return fail(OB_ERR_INVALID, "halo must be >= 1");
            if (g.topo[k] == PERIODIC && g.N[k] < g.H[k]) return fail(OB_ERR_INVALID, "periodic size smaller than halo");
        }
        for (int m = 0; m < d->n_closures; m++)
            if (d->closures[m].dynamic) {
                if (d->closures[m].kind != OB_CLOSURE_SMAGORINSKY) return fail(OB_ERR_INVALID, "a dynamic coefficient belongs to a Smagorinsky closure");
                if (d->closures[m].averaging_dims < 1 || d->closures[m].averaging_dims > 7) return fail(OB_ERR_INVALID, "DynamicSmagorinsky: averaging_dims must name at least one of the dimensions 1, 2, 3");
                if (ctx->world > 1) return fail(OB_ERR_UNSUPPORTED, "DynamicSmagorinsky on a distributed grid is not supported");
                for (int k = 0; k < 3; k++) {
                    if (g.topo[k] == FLAT) return fail(OB_ERR_UNSUPPORTED, "DynamicSmagorinsky with a Flat direction is not supported");
                    if (g.H[k] < 2) return fail(OB_ERR_INVALID, "DynamicSmagorinsky needs a halo of at least 2 cells");
                }
            }
        for (int m = 0; m < d->n_closures; m++)
            if (d->closures[m].vertically_implicit) {
                if (g.topo[2] != BOUNDED)
                    return fail(OB_ERR_INVALID, "VerticallyImplicitTimeDiscretization can only be specified on grids that are Bounded in the z-direction");
            }
        setup_field(OB_FIELD_U, 1, 0, 0, d->bcs_u);
        setup_field(OB_FIELD_V, 0, 1, 0, d->bcs_v);
        setup_field(OB_FIELD_W, 0, 0, 1, d->bcs_w);
        setup_field(OB_FIELD_PNHS, 0, 0, 0, d->bcs_p);
        if (d->has_hydrostatic_pressure) setup_field(OB_FIELD_PHY, 0, 0, 0, d->bcs_phy);
        for (int t = 0; t < ntr; t++) setup_field(OB_FIELD_TRACER0 + t, 0, 0, 0, d->bcs_tracer[t]);
        ob_bc_desc none;
        memset(&none, 0, sizeof(none));
        for (int n = 0; n < 3 + ntr; n++) {
            int lx = n == 0, ly = n == 1, lz = n == 2;
            setup_field(OB_FIELD_GN0 + n, lx, ly, lz, none);
            setup_field(OB_FIELD_GM0 + n, lx, ly, lz, none);
        }
        for (int m = 0; m < ncl; m++) {
            if (d->closures[m].kind != OB_CLOSURE_SCALAR_DIFFUSIVITY) setup_field(OB_FIELD_NUE0 + m, 0, 0, 0, d->bcs_nue[m]);
            if (d->closures[m].kind == OB_CLOSURE_AMD)
                for (int t = 0; t < ntr; t++) setup_field(OB_FIELD_KAPPAE0 + m * OB_MAX_TRACERS + t, 0, 0, 0, d->bcs_kappae[m][t]);
        }
        OB_TRY(make_solver<T>(ctx, &d->grid, &solver));
        dist = ctx->world > 1;
        if (dist && g.topo[0] != PERIODIC) return fail(OB_ERR_UNSUPPORTED, "slab-x distributed models need a Periodic x");
        if (dist && !getenv("OB_DIST_NO_IPC")) OB_TRY(setup_p2p_halo());
        return OB_OK;
    }

    int32_t bind(int32_t id, void *p) override {
        if (id < 0 || id >= 128 || !F[id].exists) return fail(OB_ERR_INVALID, "field id %d does not exist in this model", id);
        F[id].ptr = p;
        return OB_OK;
    }
    int32_t set_bc_array(int id, int side, const void *p) override {
        if (id < 0 || id >= 128 || !F[id].exists) return fail(OB_ERR_INVALID, "unknown field id %d", id);
        if (side < 0 || side >= 6) return fail(OB_ERR_INVALID, "side %d out of range", side);
        const int k = F[id].bc.kind[side];
        if (p && k != OB_BC_FLUX && k != OB_BC_VALUE && k != OB_BC_GRADIENT)
            return fail(OB_ERR_INVALID, "array-valued conditions need a Flux, Value or Gradient boundary condition on that side");
        F[id].bc_array[side] = p;
        return OB_OK;
    }
    int32_t need(int id) const {
        if (!F[id].exists || !F[id].ptr) return fail(OB_ERR_UNBOUND, "field id %d is not bound (ob_model_bind_field)", id);
        return OB_OK;
    }
    int32_t need_all() const {
        for (int id = 0; id < 128; id++)
            if (F[id].exists && !F[id].ptr) return fail(OB_ERR_UNBOUND, "field id %d is not bound (ob_model_bind_field)", id);
        return OB_OK;
    }

    // ---- halos --------------------------------------------------------------------------------------------------
    // fill_halo_regions! for a list of fields: per direction ONE launch for all fields; Bounded directions first,
    // then Periodic (boundary_condition_ordering.jl:17-46,116-142; within a class the reference order is z, y, x).
    // Distributed runs: `x_ids` (when given) replaces `ids` for the west/east exchange -- the local y / z fills of a field and
    // its exchange can then be issued at different times (x comes last in the reference order, so splitting it off is safe);
    // fields queued in `pending_x_ids` join the next exchange.
    std::vector<int> pending_x_ids;
    int32_t fill_halos(const std::vector<int> &ids, bool fill_normal, bool defer_x = false, const std::vector<int> *x_ids = nullptr) {
        PhaseScope ps(this, PH_HALO);
        // single device, triply periodic: one launch does the three periodic fills of every field (halo_periodic3_kernel)
        bool all_periodic = !dist && opt_vector && g.topo[0] == PERIODIC && g.topo[1] == PERIODIC && g.topo[2] == PERIODIC && g.N[0] >= g.H[0] &&
                            g.N[1] >= g.H[1] && g.N[2] >= g.H[2] && !ids.empty();
        for (int id : ids)
            for (int sd = 0; sd < 6; sd++) all_periodic = all_periodic && F[id].bc.kind[sd] == OB_BC_PERIODIC;
        if (all_periodic) {
            size_t pos = 0;
            while (pos < ids.size()) {
                HaloBatch<T> B;
                B.count = 0; B.dir = 0; B.N = g.N[0]; B.H = g.H[0]; B.fill_normal = 0;
                for (int k = 0; k < 3; k++) B.Hother[k] = g.H[k];
                for (; pos < ids.size() && B.count < OB_MAX_HALO_TASKS; pos++) {
                    const FieldInfo &f = F[ids[pos]];
                    OB_TRY(need(ids[pos]));
                    HaloTask<T> &t = B.t[B.count++];
                    memset(&t, 0, sizeof(t));
                    t.p = (T *)f.ptr;
                    for (int k = 0; k < 3; k++) { t.P[k] = f.P[k]; t.n[k] = f.n[k]; }
                }
                if (B.count == 0) continue;
                const FieldInfo &f0 = F[ids[0]];
                const long shell = (long)f0.P[0] * f0.P[1] * 2 * g.H[2] + (long)f0.P[0] * 2 * g.H[1] * g.N[2] + (long)2 * g.H[0] * g.N[1] * g.N[2];
                halo_periodic3_kernel<T><<<dim3(nblk(shell, 256), B.count), 256, 0, ctx->stream>>>(B, g.N[0], g.N[1], g.N[2], g.H[0], g.H[1], g.H[2]);
                launches++;
            }
            CUDA_TRY(cudaGetLastError());
            return OB_OK;
        }
        for (int pass = 0; pass < 2; pass++)
            for (int d = 2; d >= 0; d--) {
                if (g.topo[d] == FLAT) continue;
                const bool per = g.topo[d] == PERIODIC;
                if ((pass == 0) == per) continue;
                if (dist && d == 0) {
                    std::vector<int> xl = x_ids ? *x_ids : ids;
                    if (!x_ids && !pending_x_ids.empty()) {
                        for (int id : pending_x_ids) if (std::find(xl.begin(), xl.end(), id) == xl.end()) xl.push_back(id);
                        pending_x_ids.clear();
                    }
                    if (!xl.empty()) OB_TRY(exchange_x_halos(xl, defer_x));
                    continue;
                }
                size_t pos = 0;
                while (pos < ids.size()) {
                    HaloBatch<T> B;
                    B.count = 0; B.dir = d; B.N = g.N[d]; B.H = g.H[d]; B.fill_normal = fill_normal ? 1 : 0;
                    for (int k = 0; k < 3; k++) B.Hother[k] = g.H[k];
                    long maxthreads = 0;
                    for (; pos < ids.size() && B.count < OB_MAX_HALO_TASKS; pos++) {
                        const FieldInfo &f = F[ids[pos]];
                        OB_TRY(need(ids[pos]));
                        const int lo = f.bc.kind[2 * d], hi = f.bc.kind[2 * d + 1];
                        if (lo == OB_BC_NONE && hi == OB_BC_NONE) continue;
                        if ((lo == OB_BC_PERIODIC) != per || (hi == OB_BC_PERIODIC) != per)
                            return fail(OB_ERR_INVALID, "boundary condition of field %d along %d does not match the topology", ids[pos], d);
                        HaloTask<T> &t = B.t[B.count++];
                        t.p = (T *)f.ptr;
                        for (int k = 0; k < 3; k++) { t.P[k] = f.P[k]; t.n[k] = f.n[k]; }
                        t.face = f.loc[d];
                        t.bc_lo = lo; t.bc_hi = hi;
                        t.v_lo = (T)f.bc.value[2 * d]; t.v_hi = (T)f.bc.value[2 * d + 1];
                        t.a_lo = (const T *)f.bc_array[2 * d]; t.a_hi = (const T *)f.bc_array[2 * d + 1];
                        // Δ at flip(loc) at the boundary index (fill_halo_regions_value_gradient.jl:35-119)
                        auto sp = [&](int idx) -> T {
                            if (d == 0) return g.dx;
                            if (d == 1) return g.dy;
                            return f.loc[2] ? hDZC(idx) : hDZF(idx);
                        };
                        t.d_lo = sp(1);
                        t.d_hi = sp(g.N[d] + 1);
                        const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;
                        long th = per ? (long)f.P[da] * f.P[db] : (long)f.n[da] * f.n[db];
                        maxthreads = std::max(maxthreads, th);
                    }
                    if (B.count == 0) continue;
                    dim3 grid(nblk(maxthreads, 256), B.count);
                    halo_kernel<T><<<grid, 256, 0, ctx->stream>>>(B);
                    launches++;
                }
            }
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // staging buffers large enough for every field of the model in one exchange; handles all-gathered over NCCL
    int32_t setup_p2p_halo() {
        size_t cap = 0;
        for (int id = 0; id < 128; id++) if (F[id].exists) cap += (size_t)g.H[0] * F[id].P[1] * F[id].P[2];
        stage_cap = cap;
        CUDA_TRY(cudaMalloc(&d_stage, sizeof(T) * 2 * OB_HALO_SLOTS * cap));
        CUDA_TRY(cudaMalloc(&d_flags, sizeof(int) * 2 * OB_HALO_SLOTS));
        CUDA_TRY(cudaMalloc(&d_blockctr, sizeof(unsigned)));
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int) * 2 * OB_HALO_SLOTS, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(d_blockctr, 0, sizeof(unsigned), ctx->stream));
        struct Pair { cudaIpcMemHandle_t s, f; };
        Pair mine;
        const int R = ctx->world, rank = ctx->rank, west = (rank + R - 1) % R, east = (rank + 1) % R;
        int ok = 1;
        if (cudaIpcGetMemHandle(&mine.s, d_stage) != cudaSuccess || cudaIpcGetMemHandle(&mine.f, d_flags) != cudaSuccess) { cudaGetLastError(); ok = 0; memset(&mine, 0, sizeof(mine)); }
        Pair *d_all = nullptr;
        CUDA_TRY(cudaMalloc(&d_all, sizeof(Pair) * R));
        CUDA_TRY(cudaMemcpyAsync(d_all + rank, &mine, sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ncclAllGather(d_all + rank, d_all, sizeof(Pair), ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
        std::vector<Pair> all(R);
        CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(Pair) * R, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_all);
        if (ok) {
            void *ps = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&ps, all[west].s, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&pf, all[west].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            west_stage = (T *)ps; west_flags = (int *)pf;
            if (ok && east != west) {
                if (cudaIpcOpenMemHandle(&ps, all[east].s, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&pf, all[east].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
                east_stage = (T *)ps; east_flags = (int *)pf;
            } else if (ok) { east_stage = west_stage; east_flags = west_flags; }
        }
        int *d_ok = nullptr;   // every rank must take the same path
        CUDA_TRY(cudaMalloc(&d_ok, sizeof(int)));
        CUDA_TRY(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_ok);
        p2p_halo = ok != 0;
        return OB_OK;
    }
    // Staging buffers and flags form a ring of OB_HALO_SLOTS exchanges (slot = epoch mod OB_HALO_SLOTS).  An exchange is
    // a PUSH (never blocks) and a WAIT+UNPACK; with `defer` the wait is queued and issued by finish_x_halos(), so that
    // kernels which do not read x halos run while the slabs are in flight.  At most two exchanges are ever pending
    // (prognostic fields, pHY'), and a slot is reused only after this rank has completed the wait of the exchange
    // OB_HALO_SLOTS - 1 = 3 epochs later, which its neighbours can only have pushed after unpacking the slot's
    // previous contents (their pushes are stream-ordered after their earlier unpacks, two pending at most).
    struct PendingX { XHaloBatch<T> B; dim3 grid; int epoch; };
    std::vector<PendingX> pending_x;
    int32_t exchange_x_halos_p2p(const XHaloBatch<T> &B, size_t total, long maxrows, bool defer) {
        if (total > stage_cap) return fail(OB_ERR_INVALID, "halo staging buffer too small");
        if (pending_x.size() >= 2) return fail(OB_ERR_INVALID, "more than two deferred halo exchanges");
        const int epoch = ++halo_epoch, slot = epoch % OB_HALO_SLOTS;
        // layout: stage[(slot*2 + side)*cap], side 0 = west halo data, 1 = east halo data ; flags[slot*2 + side]
        auto stage = [&](T *base, int side) { return base + (size_t)(slot * 2 + side) * stage_cap; };
        dim3 grid(nblk(maxrows * g.H[0], 256), B.count);
        xhalo_push_kernel<T><<<grid, 256, 0, ctx->stream>>>(B, stage(west_stage, 1), stage(east_stage, 0), west_flags + slot * 2 + 1,
                                                            east_flags + slot * 2 + 0, d_blockctr, epoch);
        launches += 1;
        CUDA_TRY(cudaGetLastError());
        pending_x.push_back(PendingX{B, grid, epoch});
        return defer ? OB_OK : finish_x_halos();
    }
    int32_t finish_x_halos() {
        for (const PendingX &p : pending_x) {
            const int slot = p.epoch % OB_HALO_SLOTS;
            xhalo_wait_unpack_kernel<T><<<p.grid, 256, 0, ctx->stream>>>(p.B, d_stage + (size_t)(slot * 2 + 0) * stage_cap,
                                                                         d_stage + (size_t)(slot * 2 + 1) * stage_cap, d_flags + slot * 2 + 0,
                                                                         d_flags + slot * 2 + 1, p.epoch);
            launches += 1;
        }
        pending_x.clear();
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // Distributed west/east halos (halo_communication.jl:96-203; nccl_distributed.jl:209-246): ONE pack launch, one
    // grouped Send/Recv pair per side carrying every field of the batch, ONE unpack launch.  Slabs span the full
    // parent extent in y and z (OneDBuffer, communication_buffers.jl:96-114), so corners stay consistent exactly as
    // with the local periodic copy.
    int32_t exchange_x_halos(const std::vector<int> &ids, bool defer = false) {
        // more fields than one batch descriptor holds (several AMD closures x several tracers): one push/unpack epoch per
        // chunk, as fill_halos() does for the local directions
        if (ids.size() > OB_MAX_HALO_TASKS) {
            if (defer) return fail(OB_ERR_INVALID, "a deferred halo exchange is limited to %d fields", OB_MAX_HALO_TASKS);
            for (size_t pos = 0; pos < ids.size(); pos += OB_MAX_HALO_TASKS) {
                std::vector<int> part(ids.begin() + pos, ids.begin() + std::min(ids.size(), pos + (size_t)OB_MAX_HALO_TASKS));
                OB_TRY(exchange_x_halos(part, false));
            }
            return OB_OK;
        }
        XHaloBatch<T> B;
        B.count = 0; B.H = g.H[0]; B.N = g.N[0];
        size_t total = 0;
        for (int id : ids) {
            const FieldInfo &f = F[id];
            OB_TRY(need(id));
            if (f.bc.kind[0] == OB_BC_NONE && f.bc.kind[1] == OB_BC_NONE) continue;
            if (B.count >= OB_MAX_HALO_TASKS) return fail(OB_ERR_INVALID, "too many fields in one halo exchange");
            XHaloTask<T> &t = B.t[B.count++];
            t.p = (T *)f.ptr; t.Px = f.P[0]; t.rows = (long)f.P[1] * f.P[2]; t.offset = (long)total;
            total += (size_t)g.H[0] * t.rows;
        }
        if (B.count == 0) return OB_OK;
        if (p2p_halo) {
            long mr = 0;
            for (int q = 0; q < B.count; q++) mr = std::max(mr, B.t[q].rows);
            return exchange_x_halos_p2p(B, total, mr, defer);
        }
        if (4 * total > halo_buf_elems) {
            cudaFree(d_halo_buf);
            CUDA_TRY(cudaMalloc(&d_halo_buf, sizeof(T) * 4 * total));
            halo_buf_elems = 4 * total;
        }
        T *send_w = d_halo_buf, *send_e = d_halo_buf + total, *recv_w = d_halo_buf + 2 * total, *recv_e = d_halo_buf + 3 * total;
        long maxrows = 0;
        for (int q = 0; q < B.count; q++) maxrows = std::max(maxrows, B.t[q].rows);
        dim3 grid(nblk(maxrows * g.H[0], 256), B.count);
        xhalo_pack_kernel<T><<<grid, 256, 0, ctx->stream>>>(B, send_w, send_e);
        ncclComm_t comm = (ncclComm_t)ctx->comm;
        const int R = ctx->world, west = (ctx->rank + R - 1) % R, east = (ctx->rank + 1) % R;
        const size_t bytes = sizeof(T) * total;
        NCCL_TRY(ncclGroupStart());
        NCCL_TRY(ncclSend(send_w, bytes, ncclChar, west, comm, ctx->stream));   // my west interior slab -> west neighbour's east halo
        NCCL_TRY(ncclSend(send_e, bytes, ncclChar, east, comm, ctx->stream));
        NCCL_TRY(ncclRecv(recv_e, bytes, ncclChar, east, comm, ctx->stream));   // east first: matches the send order when R == 2
        NCCL_TRY(ncclRecv(recv_w, bytes, ncclChar, west, comm, ctx->stream));
        NCCL_TRY(ncclGroupEnd());
        xhalo_unpack_kernel<T><<<grid, 256, 0, ctx->stream>>>(B, recv_w, recv_e);
        launches += 3;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    int32_t fill_halo(int32_t id, int32_t fill_normal) override {
        if (id < 0 || id >= 128 || !F[id].exists) return fail(OB_ERR_INVALID, "bad field id %d", id);
        return fill_halos({id}, fill_normal != 0);
    }

    // ---- tendencies ---------------------------------------------------------------------------------------------
    TendP<T> tend_params() const {
        TendP<T> P;
        memset(&P, 0, sizeof(P));
        P.g = g;
        P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W);
        P.Gu = fld(OB_FIELD_GN0); P.Gv = fld(OB_FIELD_GN0 + 1); P.Gw = fld(OB_FIELD_GN0 + 2);
        for (int t = 0; t < ntr; t++) { P.c[t] = fld(OB_FIELD_TRACER0 + t); P.Gc[t] = fld(OB_FIELD_GN0 + 3 + t); }
        P.has_pHY = desc.has_hydrostatic_pressure;
        if (P.has_pHY) P.pHY = fld(OB_FIELD_PHY);
        P.ntr = ntr; P.ncl = ncl;
        for (int m = 0; m < ncl; m++) {
            const ob_closure_desc &c = desc.closures[m];
            ClosureD<T> &o = P.cl[m];
            o.vi = c.vertically_implicit ? 1 : 0;
            o.kind = c.kind; o.nu = (T)c.nu; o.cs = (T)c.cs; o.cb = (T)c.cb; o.lilly = c.lilly; o.Cnu = (T)c.Cnu; o.amd_has_cb = c.amd_has_cb;
            for (int t = 0; t < OB_MAXTR; t++) { o.kappa[t] = (T)c.kappa[t]; o.Pr[t] = (T)c.Pr[t]; o.Ckappa[t] = (T)c.Ckappa[t]; }
            if (c.kind != OB_CLOSURE_SCALAR_DIFFUSIVITY) P.nue[m] = fld(OB_FIELD_NUE0 + m);
            if (c.kind == OB_CLOSURE_AMD) for (int t = 0; t < ntr; t++) P.kappae[m][t] = fld(OB_FIELD_KAPPAE0 + m * OB_MAX_TRACERS + t);
        }
        P.buoy = desc.buoyancy_kind; P.ib = desc.buoyancy_tracer; P.iT = desc.temperature_tracer; P.iS = desc.salinity_tracer;
        P.grav = (T)desc.g; P.alpha = (T)desc.thermal_expansion; P.beta = (T)desc.haline_contraction;
        P.has_cor = desc.has_coriolis; P.f = (T)desc.f;
        return P;
    }

    int32_t compute_tendencies() override { return compute_tendencies_tiles(0, -1, false); }
    // x tiles [tx_lo, tx_hi) of the marching kernel (tx_hi < 0: every tile), or with `invert` every tile except those
    int32_t compute_tendencies_tiles(int tx_lo, int tx_hi, bool invert) {
        OB_TRY(need_all());
        PhaseScope ps(this, PH_TENDENCY);
        TendP<T> P = tend_params();
        const int kind = desc.advection_kind;
        const int nb = kind == OB_ADV_WENO ? (desc.advection_order + 1) / 2 : kind == OB_ADV_CENTERED ? desc.advection_order / 2 : 0;
        int nl = 0;
        cudaError_t e = launch_tendency(P, kind, nb, desc.weno_division == OB_DIV_RCP_NEWTON, opt_tendency_kernel, ctx->stream, ctx->sm_count, &nl,
                                        tx_lo, tx_hi, invert ? 1 : 0);
        if (e == cudaErrorNotSupported) return fail(OB_ERR_UNSUPPORTED, "advection scheme kind %d buffer %d has no tendency kernel", kind, nb);
        if (e != cudaSuccess) return fail(OB_ERR_CUDA, "tendency launch: %s", cudaGetErrorString(e));
        launches += nl;
        return OB_OK;
    }

    // ---- closure fields / hydrostatic pressure ---------------------------------------------------------------------
    int32_t compute_closure_fields() override {
        bool any = false;
        for (int m = 0; m < ncl; m++) any |= desc.closures[m].kind != OB_CLOSURE_SCALAR_DIFFUSIVITY;
        if (!any) return OB_OK;
        PhaseScope ps(this, PH_CLOSURE);
        TendP<T> P = tend_params();
        for (int m = 0; m < ncl; m++) {
            const int kind = desc.closures[m].kind;
            if (kind == OB_CLOSURE_SCALAR_DIFFUSIVITY) continue;
            const int bs = 128;
            dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], 1);
            if (kind == OB_CLOSURE_SMAGORINSKY && desc.closures[m].dynamic) {
                OB_TRY(dynamic_smagorinsky(m));
            } else if (kind == OB_CLOSURE_SMAGORINSKY) {
                smagorinsky_kernel<T><<<grid, bs, 0, ctx->stream>>>(P, m);
                launches++;
            } else {
                OB_TRY(amd_geometry());
                static const int minb = getenv("OB_AMD_MINB") ? atoi(getenv("OB_AMD_MINB")) : 4;
                if (minb == 5) amd_kernel<T, 5><<<grid, bs, 0, ctx->stream>>>(P, m, amd_geom);
                else if (minb == 6) amd_kernel<T, 6><<<grid, bs, 0, ctx->stream>>>(P, m, amd_geom);
                else amd_kernel<T, 4><<<grid, bs, 0, ctx->stream>>>(P, m, amd_geom);
                launches++;
            }
        }
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // DynamicSmagorinsky (dynsmag.cuh): library-owned work fields per closure -- the test-filtered velocities, Σ, Σ̄ (registered as
    // fields OB_FIELD_INTERNAL0 + 2m, + 2m + 1 with the default centre boundary conditions, so that the ordinary halo fill serves
    // them: fill_halo_regions!(Σ), fill_halo_regions!(Σ̄), dynamic_coefficient.jl:320-321), LM, MM and their averages
    static constexpr int OB_FIELD_INTERNAL0 = 112;
    struct DynWork { T *ub = nullptr, *vb = nullptr, *wb = nullptr, *Sg = nullptr, *Sb = nullptr, *LM = nullptr, *MM = nullptr, *J = nullptr; };
    DynWork dynw[OB_MAXCL];
    static DField<T> dview(const Fld<T> &f, const T *p) { DField<T> d; d.p = p; d.off = f.off; d.sy = f.sy; d.sz = f.sz; return d; }
    int32_t dynamic_smagorinsky(int m) {
        DynWork &w = dynw[m];
        const int idS = OB_FIELD_INTERNAL0 + 2 * m, idB = idS + 1;
        const int avg = desc.closures[m].averaging_dims, ax = avg & 1, ay = (avg >> 1) & 1, az = (avg >> 2) & 1;
        const long nout = (long)(ax ? 1 : g.N[0]) * (ay ? 1 : g.N[1]) * (az ? 1 : g.N[2]);
        if (!w.ub) {
            ob_bc_desc bc;
            memset(&bc, 0, sizeof(bc));
            for (int d = 0; d < 3; d++)
                for (int sd = 0; sd < 2; sd++) bc.kind[2 * d + sd] = g.topo[d] == PERIODIC ? OB_BC_PERIODIC : OB_BC_FLUX;   // NoFlux
            setup_field(idS, 0, 0, 0, bc);
            setup_field(idB, 0, 0, 0, bc);
            const size_t nu_ = F[OB_FIELD_U].count(), nv_ = F[OB_FIELD_V].count(), nw_ = F[OB_FIELD_W].count(), nc_ = F[idS].count();
            CUDA_TRY(cudaMalloc(&w.ub, sizeof(T) * nu_)); CUDA_TRY(cudaMalloc(&w.vb, sizeof(T) * nv_)); CUDA_TRY(cudaMalloc(&w.wb, sizeof(T) * nw_));
            CUDA_TRY(cudaMemsetAsync(w.ub, 0, sizeof(T) * nu_, ctx->stream)); CUDA_TRY(cudaMemsetAsync(w.vb, 0, sizeof(T) * nv_, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(w.wb, 0, sizeof(T) * nw_, ctx->stream));
            T **cq[4] = {&w.Sg, &w.Sb, &w.LM, &w.MM};
            for (int q = 0; q < 4; q++) { CUDA_TRY(cudaMalloc(cq[q], sizeof(T) * nc_)); CUDA_TRY(cudaMemsetAsync(*cq[q], 0, sizeof(T) * nc_, ctx->stream)); }
            CUDA_TRY(cudaMalloc(&w.J, sizeof(T) * 2 * nout));
            F[idS].ptr = w.Sg; F[idB].ptr = w.Sb;
        }
        DynP<T> P;
        memset(&P, 0, sizeof(P));
        for (int d = 0; d < 3; d++) { P.N[d] = g.N[d]; P.H[d] = g.H[d]; }
        P.dx = g.dx; P.dy = g.dy; P.rdx = g.rdx; P.rdy = g.rdy; P.dz = g.dz; P.rdz = g.rdz;
        P.dzc = g.dzc; P.rdzc = g.rdzc; P.rdzf = g.rdzf;
        const Fld<T> fu = fld(OB_FIELD_U), fv = fld(OB_FIELD_V), fw = fld(OB_FIELD_W), fc = fld(idS);
        P.u = dview(fu, fu.p); P.v = dview(fv, fv.p); P.w = dview(fw, fw.p);
        P.ub = dview(fu, w.ub); P.vb = dview(fv, w.vb); P.wb = dview(fw, w.wb);
        P.Sg = dview(fc, w.Sg); P.Sb = dview(fc, w.Sb); P.LM = dview(fc, w.LM); P.MM = dview(fc, w.MM);
        P.ub_w = w.ub; P.vb_w = w.vb; P.wb_w = w.wb; P.Sg_w = w.Sg; P.Sb_w = w.Sb; P.LM_w = w.LM; P.MM_w = w.MM;
        const long next = (long)(g.N[0] + 2 * g.H[0] - 2) * (g.N[1] + 2 * g.H[1] - 2) * (g.N[2] + 2 * g.H[2] - 2);
        const long ncell = (long)g.N[0] * g.N[1] * g.N[2];
        dyn_filter_kernel<T><<<nblk(next, 128), 128, 0, ctx->stream>>>(P);
        dyn_sigma_kernel<T><<<nblk(ncell, 128), 128, 0, ctx->stream>>>(P);
        launches += 2;
        CUDA_TRY(cudaGetLastError());
        OB_TRY(fill_halos({idS, idB}, true));
        dyn_lmmm_kernel<T><<<nblk(ncell, 128), 128, 0, ctx->stream>>>(P);
        dyn_average_kernel<T><<<(unsigned)nout, 256, 0, ctx->stream>>>(P, ax, ay, az, w.J);
        const Fld<T> fn = fld(OB_FIELD_NUE0 + m);
        dyn_viscosity_kernel<T><<<nblk(ncell, 128), 128, 0, ctx->stream>>>(P, ax, ay, az, w.J, (T)desc.closures[m].minimum_numerator, fn.p, fn.off, fn.sy, fn.sz);
        launches += 3;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // level-dependent metric ratios of the AMD kernel (closures.cuh: AmdGeom), evaluated once with the divisions the reference
    // performs per cell
    AmdGeom<T> amd_geom = {nullptr, 0, T(0), T(0)};
    T *d_amd_tab = nullptr;
    int32_t amd_geometry() {
        if (amd_geom.tab) return OB_OK;
        const int Nz = g.N[2], n = Nz + 2;
        std::vector<T> h(6 * (size_t)n);
        const T fx = 2 * g.dx, fy = 2 * g.dy;
        for (int k = 0; k <= Nz + 1; k++) {
            const int kk = g.topo[2] == FLAT ? 1 : k;
            const T fz = 2 * hDZC(kk);
            h[0 * n + k] = fz; h[1 * n + k] = fx / fz; h[2 * n + k] = fz / fx; h[3 * n + k] = fy / fz; h[4 * n + k] = fz / fy;
            h[5 * n + k] = 3 / (1 / (fx * fx) + 1 / (fy * fy) + 1 / (fz * fz));
        }
        CUDA_TRY(cudaMalloc(&d_amd_tab, sizeof(T) * h.size()));
        CUDA_TRY(cudaMemcpyAsync(d_amd_tab, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        amd_geom.tab = d_amd_tab; amd_geom.stride = n; amd_geom.rxy = fx / fy; amd_geom.ryx = fy / fx;
        return OB_OK;
    }
    int32_t update_hydrostatic_pressure() override { return hydrostatic_pressure(false); }
    // interior_x: only the columns i = 1 .. Nx (column-local scan: needs no x halo); the x-halo columns then come from
    // the neighbours through the halo exchange of pHY', which overwrites them anyway
    int32_t hydrostatic_pressure(bool interior_x) {
        if (!desc.has_hydrostatic_pressure || g.topo[2] == FLAT) return OB_OK;
        PhaseScope ps(this, PH_HYDRO);
        HydroP<T> P;
        memset(&P, 0, sizeof(P));
        P.g = g;
        P.pHY = fld(OB_FIELD_PHY);
        P.buoy = desc.buoyancy_kind;
        if (P.buoy == OB_BUOYANCY_TRACER) P.b = fld(OB_FIELD_TRACER0 + desc.buoyancy_tracer);
        else { P.Tt = fld(OB_FIELD_TRACER0 + desc.temperature_tracer); P.Ss = fld(OB_FIELD_TRACER0 + desc.salinity_tracer); }
        P.grav = (T)desc.g; P.alpha = (T)desc.thermal_expansion; P.beta = (T)desc.haline_contraction;
        // surface_kernel_parameters: -H+2 : N+H-1 (interleave_communication_and_computation.jl:85-94); Flat => 1:1
        P.i0 = g.topo[0] == FLAT ? 1 : -g.H[0] + 2; P.i1 = g.topo[0] == FLAT ? 1 : g.N[0] + g.H[0] - 1;
        P.j0 = g.topo[1] == FLAT ? 1 : -g.H[1] + 2; P.j1 = g.topo[1] == FLAT ? 1 : g.N[1] + g.H[1] - 1;
        if (interior_x) { P.i0 = 1; P.i1 = g.N[0]; }
        dim3 grid(nblk(P.i1 - P.i0 + 1, 64), P.j1 - P.j0 + 1);
        hydrostatic_pressure_kernel<T><<<grid, 64, 0, ctx->stream>>>(P);
        launches++;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }

    // Tiles of the marching tendency kernel that read no x halo: every x stencil of the scheme (buffer nb, and the +-1
    // neighbours of the non-advective terms) of the cells i0 .. i0+30 and of the overlap lane stays inside 1 .. Nx.
    bool interior_tiles(int &lo, int &hi) const {
        const int kind = desc.advection_kind;
        const int nb = std::max(1, kind == OB_ADV_WENO ? (desc.advection_order + 1) / 2 : kind == OB_ADV_CENTERED ? desc.advection_order / 2 : 1);
        const int ntx = (g.N[0] + OB_TILE_X - 1) / OB_TILE_X;
        lo = ntx; hi = 0;
        for (int t = 0; t < ntx; t++) {
            const int i0 = 1 + t * OB_TILE_X;
            if (i0 - nb >= 1 && i0 + OB_TILE_X + nb <= g.N[0]) { lo = std::min(lo, t); hi = std::max(hi, t + 1); }
        }
        return hi > lo;
    }
    int32_t update_state() override {
        OB_TRY(need_all());
        std::vector<int> prog = {OB_FIELD_U, OB_FIELD_V, OB_FIELD_W};
        for (int t = 0; t < ntr; t++) prog.push_back(OB_FIELD_TRACER0 + t);
        std::vector<int> aux;
        for (int m = 0; m < ncl; m++) {
            if (F[OB_FIELD_NUE0 + m].exists) aux.push_back(OB_FIELD_NUE0 + m);
            for (int t = 0; t < ntr; t++) if (F[OB_FIELD_KAPPAE0 + m * OB_MAX_TRACERS + t].exists) aux.push_back(OB_FIELD_KAPPAE0 + m * OB_MAX_TRACERS + t);
        }
        // Distributed, P2P halos, no closure fields: the x-slab pushes are issued first, the tiles that read no x halo are
        // computed while the slabs cross NVLink, the waits + unpacks come next and the edge tiles last
        // (interleave_communication_and_computation.jl:36-74, compute_nonhydrostatic_buffer_tendencies.jl)
        int lo = 0, hi = 0;
        if (dist && p2p_halo && opt_overlap && aux.empty() && opt_tendency_kernel != 1 && g.topo[0] == PERIODIC && interior_tiles(lo, hi)) {
            OB_TRY(fill_halos(prog, false, true));
            if (desc.has_hydrostatic_pressure) {
                OB_TRY(hydrostatic_pressure(true));
                OB_TRY(fill_halos({OB_FIELD_PHY}, true, true));
            }
            OB_TRY(compute_tendencies_tiles(lo, hi, false));
            {
                PhaseScope ps(this, PH_HALO);
                OB_TRY(finish_x_halos());
            }
            return compute_tendencies_tiles(lo, hi, true);   // the edge tiles, one launch
        }
        OB_TRY(fill_halos(prog, false));
        OB_TRY(compute_closure_fields());
        OB_TRY(update_hydrostatic_pressure());
        if (desc.has_hydrostatic_pressure) aux.push_back(OB_FIELD_PHY);
        if (!aux.empty()) OB_TRY(fill_halos(aux, true));
        return compute_tendencies();
    }

    // ---- time stepping ------------------------------------------------------------------------------------------------
    int32_t flux_bc_tendencies() {
        // compute_flux_bc_tendencies! (compute_nonhydrostatic_tendencies.jl:161-175): x, then y, then z, per field
        for (int d = 0; d < 3; d++)
            for (int n = 0; n < 3 + ntr; n++) {
                const int fid = n < 3 ? n : OB_FIELD_TRACER0 + (n - 3);
                const FieldInfo &f = F[fid];
                for (int side = 0; side < 2; side++) {
                    if (f.bc.kind[2 * d + side] != OB_BC_FLUX || (f.bc.value[2 * d + side] == 0.0 && !f.bc_array[2 * d + side])) continue;
                    FluxBcTask<T> t;
                    t.G = fld(OB_FIELD_GN0 + n);
                    t.dir = d; t.side = side;
                    for (int k = 0; k < 3; k++) t.loc[k] = f.loc[k];
                    t.flux = (T)f.bc.value[2 * d + side];
                    const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;
                    t.arr = (const T *)f.bc_array[2 * d + side]; t.row = f.n[da];
                    dim3 grid(nblk(g.N[da], 128), g.N[db]);
                    flux_bc_kernel<T><<<grid, 128, 0, ctx->stream>>>(g, t);
                    launches++;
                }
            }
        return OB_OK;
    }
    // single device, or distributed with a solver that hands over the west neighbour's pressure column (dist_solver.cuh)
    bool fused_substep() { return opt_fuse != 0 && (!dist || (p2p_halo && solver->enable_west_column())); }
    // what the 128-bit kernels of streaming.cuh assume of the fields one thread touches: no Flat direction, 16-byte aligned
    // base pointers, one alignment phase (same off / sy / sz modulo `v` elements; v == 2 also needs even pitches so that rows
    // j+1 and levels k+1 keep the phase), a halo row in front of the first interior row
    bool vector_ok(std::initializer_list<int> ids, int v) const {
        if (!opt_vector || g.topo[0] == FLAT || g.topo[1] == FLAT || g.topo[2] == FLAT) return false;
        if (g.H[0] < 1 || g.H[1] < 1 || g.N[0] < 8) return false;
        const Fld<T> a = fld(*ids.begin());
        for (int id : ids) {
            const Fld<T> f = fld(id);
            if (!F[id].ptr || ((uintptr_t)F[id].ptr & 15)) return false;
            if (((f.off - a.off) % v) || ((f.sy - a.sy) % v) || ((f.sz - a.sz) % v)) return false;
            if (v == 2 && ((f.sy & 1) || (f.sz & 1))) return false;
        }
        return true;
    }
    int32_t launch_update(int mode, double dt, double gamma, double zeta, double chi, bool cache) {
        PhaseScope ps(this, PH_UPDATE);
        OB_TRY(flux_bc_tendencies());
        UpdateP<T> P;
        memset(&P, 0, sizeof(P));
        P.nfields = 3 + ntr;
        for (int n = 0; n < P.nfields; n++) {
            const int fid = n < 3 ? n : OB_FIELD_TRACER0 + (n - 3);
            P.U[n] = fld(fid); P.Gn[n] = fld(OB_FIELD_GN0 + n); P.Gm[n] = fld(OB_FIELD_GM0 + n);
            for (int k = 0; k < 3; k++) P.lo[n][k] = 1;
            if (n < 3 && g.topo[n] == BOUNDED) P.lo[n][n] = 2;  // exclude_periphery (kernel_launching.jl:160-173)
        }
        for (int k = 0; k < 3; k++) P.N[k] = g.N[k];
        P.mode = mode; P.do_cache = cache ? 1 : 0;
        P.dt = (T)dt; P.gamma = (T)gamma; P.zeta = (T)zeta; P.chi = (T)chi;
        constexpr int V = Vec16<T>::V;
        bool vec = true;
        for (int n = 0; vec && n < P.nfields; n++) {
            const int fid = n < 3 ? n : OB_FIELD_TRACER0 + (n - 3);
            vec = vector_ok({fid, OB_FIELD_GN0 + n, OB_FIELD_GM0 + n}, V);
        }
        if (vec) {
            constexpr int ROWS = 2;
            const int ngx = (g.N[0] + V - 1) / V + 1;
            dim3 grid(nblk(ngx, 128) * (unsigned)((g.N[1] + ROWS - 1) / ROWS) * (unsigned)g.N[2], P.nfields);
            update_vec_kernel<T, ROWS><<<grid, 128, 0, ctx->stream>>>(P, ngx);
        } else {
            const int bs = g.N[0] >= 256 ? 256 : g.N[0] >= 128 ? 128 : g.N[0] >= 64 ? 64 : 32;
            dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], P.nfields);
            update_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
        }
        launches++;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // compute_pressure_correction! + make_pressure_correction! of one substep.  Single device: the real-part copy, the
    // correction and the p rescale are ONE kernel reading the solver output (correct_fused_kernel); distributed: the
    // reference sequence (the x neighbours of p live on other ranks and arrive through the halo exchange).
    int32_t project(double dtau) {
        if (!fused_substep()) {
            OB_TRY(compute_pressure_correction(dtau));
            return make_pressure_correction(dtau);
        }
        // distributed: the divergence reads u(Nx+1) only -- v and w travel with the next exchange of the prognostic fields
        const std::vector<int> x_u = {OB_FIELD_U}, x_none;
        OB_TRY(fill_halos({OB_FIELD_U, OB_FIELD_V, OB_FIELD_W}, true, false, dist ? &x_u : nullptr));
        const int bs = g.N[0] >= 256 ? 256 : g.N[0] >= 128 ? 128 : g.N[0] >= 64 ? 64 : 32;
        dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], 1);
        {
            PhaseScope ps(this, PH_SOURCE);
            SourceP<T> P;
            memset(&P, 0, sizeof(P));
            P.g = g; P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W);
            P.out = (T *)solver->storage();
            P.ldx = g.N[0]; P.ldxy = (long)g.N[0] * g.N[1];
            P.times_dz = solver->tridiag ? 1 : 0;
            P.cplx = solver->real_storage() ? 0 : 1;
            P.zperm = solver->z_permuted() ? 1 : 0;
            if (vector_ok({OB_FIELD_U, OB_FIELD_V, OB_FIELD_W}, 2)) {
                constexpr int ROWS = 2;
                const int ngx = g.N[0] / 2 + 1;
                source_pair_kernel<T, ROWS><<<nblk(ngx, 128) * (unsigned)((g.N[1] + ROWS - 1) / ROWS) * (unsigned)g.N[2], 128, 0, ctx->stream>>>(P, ngx);
            } else {
                source_term_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
            }
            launches++;
        }
        {
            PhaseScope ps(this, PH_SOLVE);
            int64_t before = solver->launches;
            OB_TRY(solver->solve_in_storage());
            launches += solver->launches - before;
        }
        {
            PhaseScope ps(this, PH_CORRECT);
            CorrectFusedP<T> P;
            memset(&P, 0, sizeof(P));
            P.g = g; P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W); P.p = fld(OB_FIELD_PNHS);
            P.sol = (const T *)solver->solution();
            P.ldx = solver->solution_ldx(); P.ldxy = P.ldx * g.N[1];
            P.cplx = solver->real_storage() ? 0 : 1;
            P.zperm = solver->z_permuted() ? 1 : 0;
            P.west = solver->has_west_column() ? 1 : 0;
            P.scale = (T)solver->scale();
            P.denom = std::max(std::numeric_limits<T>::epsilon(), (T)dtau);
            if (vector_ok({OB_FIELD_U, OB_FIELD_V, OB_FIELD_W, OB_FIELD_PNHS}, 2)) {
                constexpr int ROWS = 2;
                const int ngx = g.N[0] / 2 + 1;
                correct_pair_kernel<T, ROWS><<<nblk(ngx, 128) * (unsigned)((g.N[1] + ROWS - 1) / ROWS) * (unsigned)g.N[2], 128, 0, ctx->stream>>>(P, ngx);
            } else {
                correct_fused_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
            }
            launches++;
        }
        // halos of the final p: copies of interior values, as after the reference's rescale.  Distributed: nothing reads the x
        // halos of p before the next update_state!, so its slabs join that exchange instead of synchronising the ranks here
        OB_TRY(fill_halos({OB_FIELD_PNHS}, true, false, dist ? &x_none : nullptr));
        if (dist) pending_x_ids.push_back(OB_FIELD_PNHS);
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    // implicit_step! of every prognostic field after its explicit update (nonhydrostatic_rk3_substep.jl:47-56,
    // nonhydrostatic_ab2_step.jl:41-50): one launch, blockIdx.y = field (the solves are column-local and independent)
    T *d_ivd = nullptr;
    int32_t implicit_step(double dtau) {
        int nvi = 0;
        for (int m = 0; m < ncl; m++) nvi += desc.closures[m].vertically_implicit ? 1 : 0;
        if (nvi == 0) return OB_OK;
        PhaseScope ps(this, PH_UPDATE);
        IvdP<T> P;
        memset(&P, 0, sizeof(P));
        P.g = g; P.nfields = 3 + ntr; P.ntr = ntr; P.nvi = 0; P.dt = (T)dtau;
        for (int m = 0; m < ncl; m++) {
            if (!desc.closures[m].vertically_implicit) continue;
            P.nu[P.nvi] = (T)desc.closures[m].nu;
            P.kind[P.nvi] = desc.closures[m].kind;
            for (int t = 0; t < OB_MAXTR; t++) { P.kappa[P.nvi][t] = (T)desc.closures[m].kappa[t]; P.Pr[P.nvi][t] = (T)desc.closures[m].Pr[t]; }
            if (desc.closures[m].kind != OB_CLOSURE_SCALAR_DIFFUSIVITY) P.nue[P.nvi] = fld(OB_FIELD_NUE0 + m);
            if (desc.closures[m].kind == OB_CLOSURE_AMD)
                for (int t = 0; t < ntr; t++) P.kappae[P.nvi][t] = fld(OB_FIELD_KAPPAE0 + m * OB_MAX_TRACERS + t);
            P.nvi++;
        }
        for (int n = 0; n < P.nfields; n++) P.f[n] = fld(n < 3 ? n : OB_FIELD_TRACER0 + (n - 3));
        const size_t cols = (size_t)g.N[0] * g.N[1];
        if (!d_ivd) CUDA_TRY(cudaMalloc(&d_ivd, sizeof(T) * (size_t)P.nfields * (g.N[2] + 1) * cols));
        P.scratch = d_ivd;
        dim3 grid(nblk((long)cols, 128), P.nfields);
        ivd_solve_kernel<T><<<grid, 128, 0, ctx->stream>>>(P);
        launches++;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    int32_t rk3_substep(double dt, double gamma, double zeta, int has_zeta, bool cache) override {
        OB_TRY(need_all());
        OB_TRY(launch_update(has_zeta ? 1 : 0, dt, gamma, zeta, 0, cache));
        // Δτ = convert(FT, Δt * (γ + ζ)) with γ, ζ of the grid float type (runge_kutta_3.jl:186-187)
        T gz = has_zeta ? (T)((T)gamma + (T)zeta) : (T)gamma;
        double dtau = (double)(T)(dt * (double)gz);
        OB_TRY(implicit_step(dtau));
        return project(dtau);
    }
    int32_t ab2_step(double dt, double chi, bool cache) override {
        OB_TRY(need_all());
        OB_TRY(launch_update(2, dt, 0, 0, chi, cache));
        double dtau = (double)(T)dt;
        OB_TRY(implicit_step(dtau));
        return project(dtau);
    }
    int32_t cache_tendencies() override {
        OB_TRY(need_all());
        CopyP<T> P;
        memset(&P, 0, sizeof(P));
        P.nfields = 3 + ntr;
        for (int n = 0; n < P.nfields; n++) { P.dst[n] = fld(OB_FIELD_GM0 + n); P.src[n] = fld(OB_FIELD_GN0 + n); }
        for (int k = 0; k < 3; k++) P.N[k] = g.N[k];
        const int bs = 128;
        dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], P.nfields);
        cache_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
        launches++;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }

    int32_t compute_pressure_correction(double dtau) override {
        (void)dtau;
        OB_TRY(need_all());
        OB_TRY(fill_halos({OB_FIELD_U, OB_FIELD_V, OB_FIELD_W}, true));
        const int bs = g.N[0] >= 256 ? 256 : g.N[0] >= 128 ? 128 : g.N[0] >= 64 ? 64 : 32;
        dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], 1);
        {
            PhaseScope ps(this, PH_SOURCE);
            SourceP<T> P;
            memset(&P, 0, sizeof(P));
            P.g = g; P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W);
            P.out = (T *)solver->storage();
            P.ldx = g.N[0]; P.ldxy = (long)g.N[0] * g.N[1];
            P.times_dz = solver->tridiag ? 1 : 0;
            P.cplx = solver->real_storage() ? 0 : 1;
            P.zperm = solver->z_permuted() ? 1 : 0;
            if (vector_ok({OB_FIELD_U, OB_FIELD_V, OB_FIELD_W}, 2)) {
                constexpr int ROWS = 2;
                const int ngx = g.N[0] / 2 + 1;
                source_pair_kernel<T, ROWS><<<nblk(ngx, 128) * (unsigned)((g.N[1] + ROWS - 1) / ROWS) * (unsigned)g.N[2], 128, 0, ctx->stream>>>(P, ngx);
            } else {
                source_term_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
            }
            launches++;
        }
        {
            PhaseScope ps(this, PH_SOLVE);
            int64_t before = solver->launches;
            OB_TRY(solver->solve_in_storage());
            launches += solver->launches - before;
            CopyRealP<T> P;
            memset(&P, 0, sizeof(P));
            P.p = fld(OB_FIELD_PNHS);
            P.in = (const T *)solver->storage();
            P.ldx = g.N[0]; P.ldxy = (long)g.N[0] * g.N[1];
            for (int k = 0; k < 3; k++) P.N[k] = g.N[k];
            P.cplx = solver->real_storage() ? 0 : 1;
            P.zperm = solver->z_permuted() ? 1 : 0;
            P.scale = (T)solver->scale();
            copy_real_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
            launches++;
        }
        OB_TRY(fill_halos({OB_FIELD_PNHS}, true));
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }
    int32_t make_pressure_correction(double dtau) override {
        OB_TRY(need_all());
        PhaseScope ps(this, PH_CORRECT);
        CorrectP<T> P;
        memset(&P, 0, sizeof(P));
        P.g = g; P.u = fld(OB_FIELD_U); P.v = fld(OB_FIELD_V); P.w = fld(OB_FIELD_W); P.p = fld(OB_FIELD_PNHS);
        const int bs = g.N[0] >= 256 ? 256 : g.N[0] >= 128 ? 128 : g.N[0] >= 64 ? 64 : 32;
        dim3 grid(nblk(g.N[0], bs) * (unsigned)g.N[1] * (unsigned)g.N[2], 1);
        pressure_correct_kernel<T><<<grid, bs, 0, ctx->stream>>>(P);
        const long n = F[OB_FIELD_PNHS].count();
        T denom = std::max(std::numeric_limits<T>::epsilon(), (T)dtau);
        scale_kernel<T><<<nblk(n, 256), 256, 0, ctx->stream>>>((T *)F[OB_FIELD_PNHS].ptr, n, denom);
        launches += 2;
        CUDA_TRY(cudaGetLastError());
        return OB_OK;
    }

    int32_t time_step_rk3(double dt, int first) override {
        // RK3 coefficients are stored in the grid float type (runge_kutta_3.jl:66-75)
        const T g1 = (T)(8.0 / 15.0), g2 = (T)(5.0 / 12.0), g3 = (T)(3.0 / 4.0), z2 = (T)(-17.0 / 60.0), z3 = (T)(-5.0 / 12.0);
        if (first) OB_TRY(update_state());
        // cache_previous_tendencies! between the stages (runge_kutta_3.jl:125-152) is a copy G⁻ <- Gⁿ in the reference; here
        // the two tendency sets swap roles after stages 1 and 2 (the next tendencies overwrite what was G⁻) and only stage 3
        // copies, so that the step ends with Gⁿ and G⁻ where the host bound them.  Same values everywhere, 2(3+n) fewer
        // field-sized writes per step.
        const bool swap = opt_vector != 0;
        struct Reset { bool &f; ~Reset() { f = false; } } reset{gswap};   // (an error return must not leave the roles swapped)
        OB_TRY(rk3_substep(dt, g1, 0, 0, !swap));
        gswap = swap;
        OB_TRY(update_state());
        OB_TRY(rk3_substep(dt, g2, z2, 1, !swap));
        gswap = false;
        OB_TRY(update_state());
        OB_TRY(rk3_substep(dt, g3, z3, 1, true));
        OB_TRY(update_state());
        if (timing) collect();
        return OB_OK;
    }
    int32_t time_step_ab2(double dt, int euler, int first) override {
        if (first) OB_TRY(update_state());
        const double chi = euler ? -0.5 : desc.chi;
        OB_TRY(ab2_step(dt, chi, true));
        OB_TRY(update_state());
        if (timing) collect();
        return OB_OK;
    }
    int32_t advection_timescale(double *tau) override;
};

#include "diagnostics.cuh"

// ---------------------------------------------------------------------------------------------------------------
// fill_halo_regions! of ANY field array (not bound to a model): what the reference's
// fill_halo_regions!(c::OffsetArray, bcs, indices, loc, grid) does for diagnostic / user fields
// (src/BoundaryConditions/fill_halo_regions.jl:20-38).  Same kernel, same ordering as ModelT::fill_halos.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static int32_t fill_halo_array_t(ob_ctx *ctx, const ob_grid_desc *gd, void *ptr, const int32_t *loc, const ob_bc_desc *bc, const void *const *bc_arrays,
                                 int fill_normal) {
    int N[3], H[3], topo[3], n[3], P[3];
    for (int d = 0; d < 3; d++) {
        N[d] = gd->N[d]; topo[d] = gd->topology[d];
        H[d] = topo[d] == OB_FLAT ? 0 : gd->H[d];
        n[d] = N[d] + ((loc[d] && topo[d] == OB_BOUNDED) ? 1 : 0);
        P[d] = n[d] + 2 * H[d];
    }
    const T *dzf = (const T *)gd->dzf_host, *dzc = (const T *)gd->dzc_host;
    if ((dzf == nullptr) != (dzc == nullptr)) return fail(OB_ERR_INVALID, "dzf given without dzc (or the reverse)");
    for (int pass = 0; pass < 2; pass++)
        for (int d = 2; d >= 0; d--) {
            if (topo[d] == OB_FLAT) continue;
            const bool per = topo[d] == OB_PERIODIC;
            if ((pass == 0) == per) continue;
            const int lo = bc->kind[2 * d], hi = bc->kind[2 * d + 1];
            if (lo == OB_BC_NONE && hi == OB_BC_NONE) continue;
            if (lo == OB_BC_COMMUNICATION || hi == OB_BC_COMMUNICATION)
                return fail(OB_ERR_UNSUPPORTED, "communication halos are exchanged through a model (ob_fill_halo), not through ob_fill_halo_array");
            if ((lo == OB_BC_PERIODIC) != per || (hi == OB_BC_PERIODIC) != per)
                return fail(OB_ERR_INVALID, "boundary condition along %d does not match the topology", d);
            HaloBatch<T> B;
            B.count = 1; B.dir = d; B.N = N[d]; B.H = H[d]; B.fill_normal = fill_normal ? 1 : 0;
            for (int k = 0; k < 3; k++) B.Hother[k] = H[k];
            HaloTask<T> &t = B.t[0];
            t.p = (T *)ptr;
            for (int k = 0; k < 3; k++) { t.P[k] = P[k]; t.n[k] = n[k]; }
            t.face = loc[d];
            t.bc_lo = lo; t.bc_hi = hi;
            t.v_lo = (T)bc->value[2 * d]; t.v_hi = (T)bc->value[2 * d + 1];
            t.a_lo = bc_arrays ? (const T *)bc_arrays[2 * d] : nullptr;
            t.a_hi = bc_arrays ? (const T *)bc_arrays[2 * d + 1] : nullptr;
            auto sp = [&](int idx) -> T {   // Δ at flip(loc) at the boundary index (fill_halo_regions_value_gradient.jl:35-119)
                if (d == 0) return (T)gd->d[0];
                if (d == 1) return (T)gd->d[1];
                if (!dzf) return (T)gd->d[2];
                return loc[2] ? dzc[idx + H[2] - 1] : dzf[idx + H[2]];
            };
            t.d_lo = sp(1);
            t.d_hi = sp(N[d] + 1);
            const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;
            const long th = per ? (long)P[da] * P[db] : (long)n[da] * n[db];
            dim3 grid(nblk(th, 256), 1);
            halo_kernel<T><<<grid, 256, 0, ctx->stream>>>(B);
        }
    CUDA_TRY(cudaGetLastError());
    return OB_OK;
}
extern "C" int32_t ob_fill_halo_array(ob_ctx *ctx, const ob_grid_desc *grid, void *device_ptr, const int32_t *loc, const ob_bc_desc *bcs,
                                      const void *const *bc_arrays, int32_t fill_normal_flow_bcs) {
    if (!ctx || !grid || !device_ptr || !loc || !bcs) return fail(OB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (grid->float_type == OB_F64) return fill_halo_array_t<double>(ctx, grid, device_ptr, loc, bcs, bc_arrays, fill_normal_flow_bcs);
    if (grid->float_type == OB_F32) return fill_halo_array_t<float>(ctx, grid, device_ptr, loc, bcs, bc_arrays, fill_normal_flow_bcs);
    return fail(OB_ERR_INVALID, "unknown float type %d", grid->float_type);
}

// ---------------------------------------------------------------------------------------------------------------
// C ABI: model
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t ob_model_create(ob_ctx *ctx, const ob_model_desc *desc, ob_model **out) {
    if (!ctx || !desc || !out) return fail(OB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    ob_model *m = nullptr;
    int32_t st;
    if (desc->grid.float_type == OB_F64) { auto *p = new ModelT<double>(); st = p->init(ctx, desc); m = p; }
    else if (desc->grid.float_type == OB_F32) { auto *p = new ModelT<float>(); st = p->init(ctx, desc); m = p; }
    else return fail(OB_ERR_INVALID, "float_type");
    if (st != OB_OK) { delete m; return st; }
    *out = m;
    return OB_OK;
}
extern "C" int32_t ob_model_destroy(ob_model *m) {
    if (m) { cudaSetDevice(m->ctx->device); cudaStreamSynchronize(m->ctx->stream); delete m; }
    return OB_OK;
}
#define MCALL(expr)                               \
    if (!m) return fail(OB_ERR_INVALID, "null model"); \
    CUDA_TRY(cudaSetDevice(m->ctx->device));      \
    return (expr);
extern "C" int32_t ob_model_bind_field(ob_model *m, int32_t id, void *p) { MCALL(m->bind(id, p)) }
extern "C" int32_t ob_model_set_bc_array(ob_model *m, int32_t id, int32_t side, const void *p) {
    if (!m) return fail(OB_ERR_INVALID, "null model");
    MCALL(m->set_bc_array(id, side, p))
}
extern "C" int32_t ob_fill_halo(ob_model *m, int32_t id, int32_t fn) { MCALL(m->fill_halo(id, fn)) }
extern "C" int32_t ob_update_state(ob_model *m) { MCALL(m->update_state()) }
extern "C" int32_t ob_compute_tendencies(ob_model *m) { MCALL(m->compute_tendencies()) }
extern "C" int32_t ob_compute_closure_fields(ob_model *m) { MCALL(m->compute_closure_fields()) }
extern "C" int32_t ob_update_hydrostatic_pressure(ob_model *m) { MCALL(m->update_hydrostatic_pressure()) }
extern "C" int32_t ob_rk3_substep(ob_model *m, double dt, double gamma, double zeta, int32_t has_zeta) { MCALL(m->rk3_substep(dt, gamma, zeta, has_zeta, false)) }
extern "C" int32_t ob_ab2_step(ob_model *m, double dt, double chi) { MCALL(m->ab2_step(dt, chi, false)) }
extern "C" int32_t ob_cache_tendencies(ob_model *m) { MCALL(m->cache_tendencies()) }
extern "C" int32_t ob_compute_pressure_correction(ob_model *m, double dtau) { MCALL(m->compute_pressure_correction(dtau)) }
extern "C" int32_t ob_make_pressure_correction(ob_model *m, double dtau) { MCALL(m->make_pressure_correction(dtau)) }
extern "C" int32_t ob_time_step_rk3(ob_model *m, double dt, int32_t first) { MCALL(m->time_step_rk3(dt, first)) }
extern "C" int32_t ob_time_step_ab2(ob_model *m, double dt, int32_t euler, int32_t first) { MCALL(m->time_step_ab2(dt, euler, first)) }
extern "C" int32_t ob_cell_advection_timescale(ob_model *m, double *tau) { MCALL(m->advection_timescale(tau)) }
extern "C" int32_t ob_model_set_option(ob_model *m, int32_t option, int32_t value) {
    if (!m) return fail(OB_ERR_INVALID, "null model");
    switch (option) {
        case OB_OPT_TENDENCY_KERNEL: m->opt_tendency_kernel = value; return OB_OK;
        case OB_OPT_FUSE_PROJECTION: m->opt_fuse = value; return OB_OK;
        case OB_OPT_OVERLAP_HALO: m->opt_overlap = value; return OB_OK;
        case OB_OPT_VECTOR_STREAMS: m->opt_vector = value; return OB_OK;
    }
    return fail(OB_ERR_INVALID, "unknown option %d", option);
}
extern "C" int32_t ob_launch_count(ob_model *m, int64_t *n) {
    if (!m || !n) return fail(OB_ERR_INVALID, "null model or output pointer");
    *n = m->launches;
    return OB_OK;
}
extern "C" int32_t ob_enable_timing(ob_model *m, int32_t e) {
    if (!m) return fail(OB_ERR_INVALID, "null model");
    m->collect();
    m->timing = e != 0;
    return OB_OK;
}
extern "C" int32_t ob_phase_count(int32_t *n) { *n = PH_COUNT; return OB_OK; }
extern "C" const char *ob_phase_name(int32_t p) { return (p >= 0 && p < PH_COUNT) ? PHASE_NAMES[p] : ""; }
extern "C" int32_t ob_phase_time_ms(ob_model *m, int32_t p, double *ms, int64_t *calls) {
    if (p < 0 || p >= PH_COUNT) return fail(OB_ERR_INVALID, "phase");
    m->collect();
    *ms = m->phase_ms[p];
    *calls = m->phase_calls[p];
    return OB_OK;
}
extern "C" int32_t ob_reset_timing(ob_model *m) {
    if (!m) return fail(OB_ERR_INVALID, "null model");
    m->collect();
    for (int p = 0; p < PH_COUNT; p++) { m->phase_ms[p] = 0; m->phase_calls[p] = 0; }
    return OB_OK;
}

#include "dist.cuh"
