// kernels.cuh -- streaming kernels of the time step: halo fills (K5/K6), flux-BC tendencies (K7), hydrostatic
// pressure scan (K8), RK3/AB2 update fused with the G⁻ <- Gⁿ cache (K3+K4), Poisson source term (K10),
// pressure correction (K16).  All are HBM-bound, coalesced along x, batched over fields so that one launch
// serves every prognostic field (the reference launches one kernel per field: SURVEY.md §2a).
#pragma once
#include "common.cuh"

namespace ob {

// ------------------------------------------------------------------------------------------------------------
// Halo fills.  Bit-exact restatement of fill_halo_regions_periodic.jl:5-27, _flux.jl:9-27,
// _value_gradient.jl:7-119, _normal_flow.jl:2-27 with the reference's launch extents
// (periodic: full parent extent of the tangential dims; others: interior extent, one halo cell).
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct HaloTask {
    T *p;          // parent array
    int P[3];      // parent sizes
    int n[3];      // interior sizes (N or N+1)
    int face;      // field is Face-located along `dir`
    int bc_lo, bc_hi;
    T v_lo, v_hi;  // constant value / gradient
    const T *a_lo, *a_hi;  // array-valued condition over the interior extent of the tangential dims (n[da] fastest), or nullptr
    T d_lo, d_hi;  // Δ at the boundary (flipped location) for Value/Gradient
};
#define OB_MAX_HALO_TASKS 24
#define OB_HALO_SLOTS 4   // ring depth of the P2P halo staging buffers / flags (ocean_b200.cu: exchange_x_halos_p2p)
template <typename T>
struct HaloBatch {
    HaloTask<T> t[OB_MAX_HALO_TASKS];
    int count, dir, N, H, fill_normal;
    int Hother[3];
};

template <typename T>
__global__ void __launch_bounds__(256) halo_kernel(const __grid_constant__ HaloBatch<T> B) {
    const HaloTask<T> &t = B.t[blockIdx.y];
    const int d = B.dir, N = B.N, H = B.H;
    const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;  // tangential dims, da the faster one
    const long sx = 1, sy = t.P[0], sz = (long)t.P[0] * t.P[1];
    const long sd = d == 0 ? sx : d == 1 ? sy : sz;
    const long sa = da == 0 ? sx : sy, sb = db == 1 ? sy : sz;
    const bool periodic = t.bc_lo == BC_PERIODIC;
    const int A = periodic ? t.P[da] : t.n[da];
    const int Bn = periodic ? t.P[db] : t.n[db];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long)A * Bn) return;
    const int a = (int)(tid % A), b = (int)(tid / A);
    T *q;
    if (periodic) {
        q = t.p + a * sa + b * sb;
#pragma unroll 1
        for (int h = 0; h < H; h++) {
            q[h * sd] = q[(N + h) * sd];          // parent[i] = parent[N+i]
            q[(N + H + h) * sd] = q[(H + h) * sd];  // parent[N+H+i] = parent[H+i]
        }
        return;
    }
    q = t.p + (a + B.Hother[da]) * sa + (b + B.Hother[db]) * sb;  // interior window of the tangential dims
    // getbc(condition::AbstractArray, i, j, ...) = condition[i, j] (boundary_condition.jl)
    const T v_lo = t.a_lo ? __ldg(t.a_lo + a + (long)b * A) : t.v_lo;
    const T v_hi = t.a_hi ? __ldg(t.a_hi + a + (long)b * A) : t.v_hi;
    // logical index l along d -> parent index l - 1 + H
#define AT(l) q[((l) - 1 + H) * sd]
    // low side
    switch (t.bc_lo) {
        case BC_FLUX: AT(0) = AT(1); break;
        case BC_IMPENETRABLE: if (B.fill_normal) AT(1) = T(0); break;
        case BC_GRADIENT: { T c = AT(1); AT(0) = add_rn(c, mul_rn(v_lo, -t.d_lo)); } break;
        case BC_VALUE: { T c = AT(1); T g = (c - v_lo) / (t.d_lo / 2); AT(0) = add_rn(c, mul_rn(g, -t.d_lo)); } break;
        default: break;
    }
    switch (t.bc_hi) {
        case BC_FLUX: AT(N + 1) = AT(N); break;
        case BC_IMPENETRABLE: if (B.fill_normal) AT(N + 1) = T(0); break;
        case BC_GRADIENT: { T c = AT(N); AT(N + 1) = add_rn(c, mul_rn(v_hi, t.d_hi)); } break;
        case BC_VALUE: { T c = AT(N); T g = (v_hi - c) / (t.d_hi / 2); AT(N + 1) = add_rn(c, mul_rn(g, t.d_hi)); } break;
        default: break;
    }
#undef AT
}

// Triply periodic fields: the three periodic fills (z, then y, then x, each over the full parent extent of the other two
// directions: fill_halo_regions_periodic.jl:5-27) in ONE launch.  After the sequence every halo cell holds the interior value
// its indices wrap to, so each thread of the shell copies that value directly -- bit-identical, and two launches fewer per fill.
// The shell is enumerated as three disjoint slabs: k in the halo (full x-y planes), j in the halo (interior k), i in the halo
// (interior j, k); x is the fastest index of each.
template <typename T>
__global__ void __launch_bounds__(256) halo_periodic3_kernel(const __grid_constant__ HaloBatch<T> B, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz) {
    const HaloTask<T> &t = B.t[blockIdx.y];
    const int Px = t.P[0], Py = t.P[1];
    const long nZ = (long)Px * Py * 2 * Hz, nY = (long)Px * 2 * Hy * Nz, nX = (long)2 * Hx * Ny * Nz;
    long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int i, j, k;   // 0-based parent indices of the halo cell
    if (q < nZ) {
        i = (int)(q % Px); j = (int)((q / Px) % Py);
        const int h = (int)(q / ((long)Px * Py));
        k = h < Hz ? h : Nz + h;
    } else if (q < nZ + nY) {
        q -= nZ;
        i = (int)(q % Px);
        const int h = (int)((q / Px) % (2 * Hy));
        j = h < Hy ? h : Ny + h;
        k = Hz + (int)(q / ((long)Px * 2 * Hy));
    } else if (q < nZ + nY + nX) {
        q -= nZ + nY;
        const int h = (int)(q % (2 * Hx));
        i = h < Hx ? h : Nx + h;
        j = Hy + (int)((q / (2 * Hx)) % Ny);
        k = Hz + (int)(q / ((long)2 * Hx * Ny));
    } else {
        return;
    }
    const int si = i < Hx ? i + Nx : i >= Hx + Nx ? i - Nx : i;
    const int sj = j < Hy ? j + Ny : j >= Hy + Ny ? j - Ny : j;
    const int sk = k < Hz ? k + Nz : k >= Hz + Nz ? k - Nz : k;
    const long sy = Px, sz = (long)Px * Py;
    t.p[i + j * sy + k * sz] = t.p[si + sj * sy + sk * sz];
}

// Distributed west/east halo slabs: pack the H interior columns next to each x boundary of every field of the batch
// into contiguous buffers, and unpack the neighbours' slabs into the halos.  Integer index work: bit-exact.
template <typename T>
struct XHaloTask {
    T *p;
    int Px;       // parent x size
    long rows;    // Py * Pz
    long offset;  // element offset of this field inside the slab buffers
};
template <typename T>
struct XHaloBatch {
    XHaloTask<T> t[OB_MAX_HALO_TASKS];
    int count, H, N;
};
template <typename T>
__global__ void __launch_bounds__(256) xhalo_pack_kernel(const __grid_constant__ XHaloBatch<T> B, T *__restrict__ send_w, T *__restrict__ send_e) {
    const XHaloTask<T> &t = B.t[blockIdx.y];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= t.rows * B.H) return;
    const int h = (int)(tid % B.H);
    const long row = tid / B.H;
    const T *q = t.p + row * t.Px;
    send_w[t.offset + tid] = q[B.H + h];   // logical i = 1 .. H
    send_e[t.offset + tid] = q[B.N + h];   // logical i = N-H+1 .. N
}
template <typename T>
__global__ void __launch_bounds__(256) xhalo_unpack_kernel(const __grid_constant__ XHaloBatch<T> B, const T *__restrict__ recv_w, const T *__restrict__ recv_e) {
    const XHaloTask<T> &t = B.t[blockIdx.y];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= t.rows * B.H) return;
    const int h = (int)(tid % B.H);
    const long row = tid / B.H;
    T *q = t.p + row * t.Px;
    q[h] = recv_w[t.offset + tid];               // west halo <- west neighbour's east interior slab
    q[B.N + B.H + h] = recv_e[t.offset + tid];   // east halo <- east neighbour's west interior slab
}

// Peer-to-peer form of the same exchange (NVLink, CUDA-IPC mapped peer buffers; no NCCL on the data path):
//   push:   every rank stores its west / east interior slabs DIRECTLY into the staging buffer of its west / east
//           neighbour, then -- once the last block has finished (system-scope fences + a block counter) -- publishes the
//           exchange's epoch number into the neighbour's flag with a release store;
//   unpack: waits (acquire loads) until both of its own flags carry the epoch, then copies the staged slabs into the halos.
// Staging buffers and flags form a ring of OB_HALO_SLOTS exchanges indexed by the epoch: with at most two exchanges
// pending per rank (pushes issued, waits deferred behind the interior tendency tiles) a neighbour can never overwrite
// a slot this rank has not unpacked yet (argument at exchange_x_halos_p2p).
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int *p) { int v; asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
template <typename T>
__global__ void __launch_bounds__(256) xhalo_push_kernel(const __grid_constant__ XHaloBatch<T> B, T *__restrict__ west_nbr_recv_e,
                                                         T *__restrict__ east_nbr_recv_w, int *west_nbr_flag_e, int *east_nbr_flag_w,
                                                         unsigned *block_counter, int epoch) {
    const XHaloTask<T> &t = B.t[blockIdx.y];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < t.rows * B.H) {
        const int h = (int)(tid % B.H);
        const long row = tid / B.H;
        const T *q = t.p + row * t.Px;
        west_nbr_recv_e[t.offset + tid] = q[B.H + h];   // my west interior slab -> west neighbour's east halo
        east_nbr_recv_w[t.offset + tid] = q[B.N + h];   // my east interior slab -> east neighbour's west halo
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(block_counter, 1u) == total - 1) {
            *block_counter = 0;
            __threadfence_system();
            st_release_sys(west_nbr_flag_e, epoch);
            st_release_sys(east_nbr_flag_w, epoch);
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(256) xhalo_wait_unpack_kernel(const __grid_constant__ XHaloBatch<T> B, const T *__restrict__ recv_w,
                                                                const T *__restrict__ recv_e, const int *flag_w, const int *flag_e, int epoch) {
    if (threadIdx.x == 0) {
        while (ld_acquire_sys(flag_w) < epoch) {}
        while (ld_acquire_sys(flag_e) < epoch) {}
    }
    __syncthreads();
    const XHaloTask<T> &t = B.t[blockIdx.y];
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= t.rows * B.H) return;
    const int h = (int)(tid % B.H);
    const long row = tid / B.H;
    T *q = t.p + row * t.Px;
    q[h] = __ldcv(recv_w + t.offset + tid);
    q[B.N + B.H + h] = __ldcv(recv_e + t.offset + tid);
}

// ------------------------------------------------------------------------------------------------------------
// implicit_step! of every prognostic field (vertically_implicit_diffusion_solver.jl:60-136,196-225): the tridiagonal
// system (1 - Δτ ∂z κ ∂z) ϕ = ϕ★ of each column, coefficients from the ivd_* functions (periphery-aware), solved in
// place by the sweep of solve_batched_tridiagonal_system_z! (batched_tridiagonal_solver.jl:211-243).  One thread per
// column, x fastest (coalesced); blockIdx.y = field.  inactive_cell / inactive_node / peripheral_node:
// src/Grids/inactive_node.jl:43-165.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct IvdP {
    GridD<T> g;
    int nfields, ntr, nvi;
    Fld<T> f[3 + OB_MAXTR];
    T nu[OB_MAXCL];                 // the vertically-implicit closures only, in closure order
    T kappa[OB_MAXCL][OB_MAXTR];
    // eddy-viscosity closures (Smagorinsky, AMD) with VerticallyImplicitTimeDiscretization: the coefficients are nu_e / kappa_e
    // interpolated to the node (abstract_scalar_diffusivity_closure.jl:137-151, 330-351)
    int kind[OB_MAXCL];
    T Pr[OB_MAXCL][OB_MAXTR];
    Fld<T> nue[OB_MAXCL];
    Fld<T> kappae[OB_MAXCL][OB_MAXTR];
    T dt;
    T *scratch;                     // nfields x (Nz+1) x Ny x Nx
};
template <typename T>
struct IvdCol {
    const IvdP<T> &P;
    int which, lx, ly, lz, i, j;
    __device__ __forceinline__ bool inactive_cell(int a, int b, int c) const {
        const GridD<T> &g = P.g;
        return ((g.topo[0] == BOUNDED) & ((a < 1) | (a > g.N[0]))) | ((g.topo[1] == BOUNDED) & ((b < 1) | (b > g.N[1]))) |
               ((g.topo[2] == BOUNDED) & ((c < 1) | (c > g.N[2])));
    }
    // any: peripheral_node (OR over the cells around the node); !any: inactive_node (AND)
    __device__ __forceinline__ bool node(int k, int fz, bool any) const {
        bool r = !any;
        for (int a = 0; a <= lx; a++)
            for (int b = 0; b <= ly; b++)
                for (int c = 0; c <= fz; c++) {
                    const bool v = inactive_cell(i - a, j - b, k - c);
                    r = any ? (r | v) : (r & v);
                }
        return r;
    }
    // two-point interpolation of a ccc array to the face below / west / south (interpolation_operators.jl:8-28), nested as the
    // reference nests them: ℑxzᶠᵃᶠ = ℑzᵃᵃᶠ(ℑxᶠᵃᵃ), ℑyzᵃᶠᶠ = ℑzᵃᵃᶠ(ℑyᵃᶠᵃ); Flat => identity
    __device__ __forceinline__ T i1(const Fld<T> &f, int d, int a, int b, int c) const {
        if (P.g.topo[d] == FLAT) return f.ld(a, b, c);
        return T(0.5) * (f.ld(a - (d == 0), b - (d == 1), c - (d == 2)) + f.ld(a, b, c));
    }
    __device__ __forceinline__ T i2z(const Fld<T> &f, int d1, int a, int b, int c) const {
        if (P.g.topo[2] == FLAT) return i1(f, d1, a, b, c);
        return T(0.5) * (i1(f, d1, a, b, c - 1) + i1(f, d1, a, b, c));
    }
    // νzᶠᶜᶠ / νzᶜᶠᶠ / νzᶜᶜᶜ / κzᶜᶜᶠ of closure m at level kk
    __device__ __forceinline__ T coef(int m, int kk) const {
        if (P.kind[m] == CL_SCALAR) return which < 3 ? P.nu[m] : P.kappa[m][which - 3];
        if (which == 0) return i2z(P.nue[m], 0, i, j, kk);
        if (which == 1) return i2z(P.nue[m], 1, i, j, kk);
        if (which == 2) return P.nue[m].ld(i, j, kk);
        if (P.kind[m] == CL_SMAG) return i1(P.nue[m], 2, i, j, kk) / P.Pr[m][which - 3];
        return i1(P.kappae[m][which - 3], 2, i, j, kk);
    }
    __device__ __forceinline__ T upper(int k) const {
        T sum = 0;
        for (int m = 0; m < P.nvi; m++) {
            T d;
            if (!lz) {
                const T kap = node(k + 1, 1, false) ? T(0) : coef(m, k + 1);
                d = -P.dt * kap * (P.g.rdzC(k) * P.g.rdzF(k + 1));
                if (node(k + 1, 1, true)) d = 0;
            } else {
                const T nu = node(k, 0, false) ? T(0) : coef(m, k);
                d = -P.dt * nu * (P.g.rdzC(k) * P.g.rdzF(k));
                if (node(k, 0, true)) d = 0;
            }
            sum = m == 0 ? d : sum + d;
        }
        return sum;
    }
    __device__ __forceinline__ T lower(int kk) const {
        T sum = 0;
        for (int m = 0; m < P.nvi; m++) {
            T d;
            if (!lz) {
                const int k = kk + 1;
                const T kap = node(k, 1, false) ? T(0) : coef(m, k);
                d = -P.dt * kap * (P.g.rdzC(k) * P.g.rdzF(k));
            } else {
                const int kp = kk + 2;
                const T nu = node(kp - 1, 0, false) ? T(0) : coef(m, kp - 1);
                d = -P.dt * nu * (P.g.rdzC(kp) * P.g.rdzF(kp - 1));
            }
            if (node(kk, 0, true)) d = 0;
            sum = m == 0 ? d : sum + d;
        }
        return sum;
    }
    __device__ __forceinline__ T diag(int k) const { return T(1) - P.dt * T(0) - upper(k) - lower(k - 1); }
};
template <typename T>
__global__ void __launch_bounds__(128) ivd_solve_kernel(const __grid_constant__ IvdP<T> P) {
    const int n = blockIdx.y;
    const int Nx = P.g.N[0], Ny = P.g.N[1], Nz = P.g.N[2];
    const long col = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= (long)Nx * Ny) return;
    const int i = 1 + (int)(col % Nx), j = 1 + (int)(col / Nx);
    const int which = n < 3 ? n : 3 + (n - 3);
    const IvdCol<T> C{P, which, which == 0, which == 1, which == 2, i, j};
    const Fld<T> &f = P.f[n];
    T *t = P.scratch + (size_t)n * (Nz + 1) * Ny * Nx + col;   // t[k] at t[k * Nx*Ny]
    const size_t ts = (size_t)Nx * Ny;
    const T tiny = 10 * (sizeof(T) == 8 ? T(2.220446049250313e-16) : T(1.1920929e-07));
    T beta = C.diag(1);
    T prev = f(i, j, 1) / beta;
    f(i, j, 1) = prev;
    for (int k = 2; k <= Nz; k++) {
        const T cm = C.upper(k - 1), bk = C.diag(k), am = C.lower(k - 1);
        const T tk = cm / beta;
        t[k * ts] = tk;
        beta = sub_rn(bk, mul_rn(am, tk));
        const T fk = f(i, j, k);
        const T cand = sub_rn(fk, mul_rn(am, prev)) / beta;
        prev = fabs(beta) > tiny ? cand : fk;
        f(i, j, k) = prev;
    }
    for (int k = Nz - 1; k >= 1; k--) {
        prev = sub_rn(f(i, j, k), mul_rn(t[(k + 1) * ts], prev));
        f(i, j, k) = prev;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K7: constant-flux boundary contributions (compute_flux_bcs.jl:113-162): Gc[1] += flux*A/V ; Gc[N] -= flux*A/V
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct FluxBcTask {
    Fld<T> G;
    int dir, side;   // side 0 = low (+=), 1 = high (-=)
    int loc[3];      // 1 = face, 0 = center
    T flux;
    const T *arr;    // array-valued flux over the field's tangential interior extent (row length `row`), or nullptr
    int row;
};
template <typename T>
__global__ void flux_bc_kernel(GridD<T> g, FluxBcTask<T> t) {
    const int d = t.dir;
    const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (a >= g.N[da] || b >= g.N[db]) return;
    int idx[3];
    idx[da] = a + 1; idx[db] = b + 1;
    const int n_int = t.side == 0 ? 1 : g.N[d];
    const int n_face = t.side == 0 ? 1 : g.N[d] + 1;
    idx[d] = n_int;
    auto sp = [&](int dd, int face, int l) -> T {
        if (dd == 0) return g.dx;
        if (dd == 1) return g.dy;
        return face ? g.dzF(l) : g.dzC(l);
    };
    // area normal to d: tangential spacings at the field's locations (z index = own index if tangential)
    T sa = sp(da, t.loc[da], idx[da]), sb = sp(db, t.loc[db], idx[db]);
    T sn = sp(d, t.loc[d], n_int);
    (void)n_face;
    T area = sa * sb;  // Ax = Δy*Δz ; Ay = Δx*Δz ; Az = Δx*Δy  (da < db always)
    T vol = d == 0 ? (sn * sa) * sb : d == 1 ? (sa * sn) * sb : (sa * sb) * sn;  // V = (Δx*Δy)*Δz
    const T flux = t.arr ? __ldg(t.arr + a + (long)b * t.row) : t.flux;
    T term = flux * area / vol;
    T &G = t.G(idx[0], idx[1], idx[2]);
    G = t.side == 0 ? G + term : G - term;
}

// ------------------------------------------------------------------------------------------------------------
// K8: hydrostatic pressure anomaly (update_hydrostatic_pressure.jl:11-39), columns i,j in (-H+2 : N+H-1)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct HydroP {
    GridD<T> g;
    Fld<T> pHY, b, Tt, Ss;
    int buoy;
    T grav, alpha, beta;
    int i0, i1, j0, j1;
};
template <typename T>
__global__ void hydrostatic_pressure_kernel(HydroP<T> P) {
    const int i = P.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = P.j0 + blockIdx.y;
    if (i > P.i1 || j > P.j1) return;
    const int Nz = P.g.N[2];
    auto bp = [&](int k) -> T {
        if (P.buoy == BUOY_TRACER) return P.b.ld(i, j, k);
        return P.grav * (P.alpha * P.Tt.ld(i, j, k) - P.beta * P.Ss.ld(i, j, k));
    };
    T bk = bp(Nz + 1), bkm = bp(Nz);
    T pk = -(T(0.5) * (bkm + bk)) * P.g.dzF(Nz + 1);
    P.pHY(i, j, Nz) = pk;
    for (int k = Nz - 1; k >= 1; k--) {
        bk = bkm;          // b[k+1]
        bkm = bp(k);       // b[k]
        pk = pk - (T(0.5) * (bkm + bk)) * P.g.dzF(k + 1);
        P.pHY(i, j, k) = pk;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K3 + K4: RK3 / AB2 update of every prognostic field, fused with the G⁻ <- Gⁿ cache
// (runge_kutta_3.jl:196-204, quasi_adams_bashforth_2.jl:134-147, cache_nonhydrostatic_tendencies.jl:8-31).
// The update skips boundary-normal faces of u, v, w (exclude_periphery); the cache covers 1..N.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct UpdateP {
    Fld<T> U[3 + OB_MAXTR], Gn[3 + OB_MAXTR], Gm[3 + OB_MAXTR];
    int lo[3 + OB_MAXTR][3];  // first updated index per dim (2 for the wall-normal component on Bounded)
    int N[3];
    int nfields;
    int mode;       // 0: rk3 first stage, 1: rk3 with zeta, 2: ab2
    int do_cache;   // also write G⁻ = Gⁿ
    T dt, gamma, zeta, chi;
};
// The update of one value, rounded operation by operation as the reference's kernels are (no @muladd there, and Julia does not
// contract): shared by update_kernel and its 128-bit form so that the two agree bit for bit whatever the compiler would fuse.
//   mode 0: U += Δt γ¹ G¹ ((Δt γ¹) first);  mode 1: U += Δt (γ Gⁿ + ζ G⁻);  mode 2: U += Δt ((3/2 + χ) Gⁿ - (1/2 + χ) G⁻ not_euler)
template <typename T>
__device__ __forceinline__ T updated_value(const UpdateP<T> &P, T u, T gn, T gm) {
    if (P.mode == 0) return add_rn(u, mul_rn(mul_rn(P.dt, P.gamma), gn));
    if (P.mode == 1) return add_rn(u, mul_rn(P.dt, add_rn(mul_rn(P.gamma, gn), mul_rn(P.zeta, gm))));
    const T a = T(1.5) + P.chi, b = T(0.5) + P.chi;
    const bool not_euler = P.chi != T(-0.5);
    const T G = sub_rn(mul_rn(a, gn), not_euler ? mul_rn(b, gm) : T(0));
    return add_rn(u, mul_rn(P.dt, G));
}
template <typename T>
__global__ void __launch_bounds__(256) update_kernel(const __grid_constant__ UpdateP<T> P) {
    const int f = blockIdx.y;
    int i, j, k;
    if (!cell_from_block(P.N[0], P.N[1], i, j, k)) return;
    const long ig = P.Gn[f].idx(i, j, k);
    const T gn = P.Gn[f].p[ig];
    const long im = P.Gm[f].idx(i, j, k);
    const bool active = (i >= P.lo[f][0]) & (j >= P.lo[f][1]) & (k >= P.lo[f][2]);
    if (active) {
        T &u = P.U[f](i, j, k);
        const bool need_gm = P.mode == 1 || (P.mode == 2 && P.chi != T(-0.5));
        u = updated_value(P, u, gn, need_gm ? P.Gm[f].p[im] : T(0));
    }
    if (P.do_cache) P.Gm[f].p[im] = gn;
}

template <typename T>
struct CopyP {
    Fld<T> dst[3 + OB_MAXTR], src[3 + OB_MAXTR];
    int N[3], nfields;
};
template <typename T>
__global__ void cache_kernel(const __grid_constant__ CopyP<T> P) {
    const int f = blockIdx.y;
    int i, j, k;
    if (!cell_from_block(P.N[0], P.N[1], i, j, k)) return;
    P.dst[f](i, j, k) = P.src[f](i, j, k);
}

// ------------------------------------------------------------------------------------------------------------
// K10: Poisson source term rhs = divᶜᶜᶜ(U★) (solve_for_pressure.jl:12-18), × Δzᶜ for the tridiagonal solver
// (:36-42).  Written as REAL numbers into the solver's input (the reference writes complex storage).
// `out` has logical layout (Nx,Ny,Nz) with leading dimension ldx (padded for in-place real-to-complex FFTs).
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct SourceP {
    GridD<T> g;
    Fld<T> u, v, w;
    T *out;
    long ldx, ldxy;
    int times_dz;
    int cplx;  // write complex (re, 0) pairs instead of reals
    int zperm; // write level k at the Makhoul-permuted position (real DCT path): even k-1 -> (k-1)/2, odd -> Nz-1-(k-2)/2
};
// divᶜᶜᶜ of one cell from its six face values, every product and sum rounded separately (divergence_operators.jl has no @muladd)
template <typename T>
__device__ __forceinline__ T source_value(const SourceP<T> &P, T Ax, T Ay, T Az, T Vi, T dzc, T u0, T u1, T v0, T v1, T w0, T w1) {
    const T ddx = P.g.topo[0] == FLAT ? T(0) : sub_rn(mul_rn(Ax, u1), mul_rn(Ax, u0));
    const T ddy = P.g.topo[1] == FLAT ? T(0) : sub_rn(mul_rn(Ay, v1), mul_rn(Ay, v0));
    const T ddz = P.g.topo[2] == FLAT ? T(0) : sub_rn(mul_rn(Az, w1), mul_rn(Az, w0));
    T div = mul_rn(Vi, add_rn(add_rn(ddx, ddy), ddz));
    if (P.times_dz) div = mul_rn(dzc, div);
    return div;
}
template <typename T>
__global__ void __launch_bounds__(256) source_term_kernel(const __grid_constant__ SourceP<T> P) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    const T dzc = P.g.dzC(k);
    const T Ax = P.g.dy * dzc, Ay = P.g.dx * dzc, Az = P.g.dx * P.g.dy;
    // (Flat directions: the index stays put -- there is no halo to read -- and source_value() drops the term)
    const int fx = P.g.topo[0] == FLAT ? 0 : 1, fy = P.g.topo[1] == FLAT ? 0 : 1, fz = P.g.topo[2] == FLAT ? 0 : 1;
    const T div = source_value(P, Ax, Ay, Az, P.g.rVc(k), dzc, P.u.ld(i, j, k), P.u.ld(i + fx, j, k), P.v.ld(i, j, k), P.v.ld(i, j + fy, k),
                               P.w.ld(i, j, k), P.w.ld(i, j, k + fz));
    const int kk = P.zperm ? makhoul_index(k - 1, P.g.N[2]) : k - 1;
    const long o = (i - 1) + (j - 1) * P.ldx + (long)kk * P.ldxy;
    if (P.cplx) { P.out[2 * o] = div; P.out[2 * o + 1] = T(0); }
    else P.out[o] = div;
}

// K15: p <- real solution (copy_real_component!, fft_based_poisson_solver.jl:128-136), with a scale factor
// that carries the unnormalised cuFFT inverse (1/N per transformed dimension).
template <typename T>
struct CopyRealP {
    Fld<T> p;
    const T *in;
    long ldx, ldxy;
    int N[3];
    int cplx;
    int zperm;
    T scale;
};
template <typename T>
__global__ void __launch_bounds__(256) copy_real_kernel(const __grid_constant__ CopyRealP<T> P) {
    int i, j, k;
    if (!cell_from_block(P.N[0], P.N[1], i, j, k)) return;
    const int kk = P.zperm ? makhoul_index(k - 1, P.N[2]) : k - 1;
    const long o = (i - 1) + (j - 1) * P.ldx + (long)kk * P.ldxy;
    P.p(i, j, k) = (P.cplx ? P.in[2 * o] : P.in[o]) * P.scale;
}

// K16: u -= ∂x(pΔτ), v -= ∂y(pΔτ), w -= ∂z(pΔτ) over :xyz (pressure_correction.jl:67-73)
template <typename T>
struct CorrectP {
    GridD<T> g;
    Fld<T> u, v, w, p;
};
template <typename T>
__global__ void __launch_bounds__(256) pressure_correct_kernel(const __grid_constant__ CorrectP<T> P) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    const T pc = P.p.ld(i, j, k);
    if (P.g.topo[0] != FLAT) P.u(i, j, k) = sub_rn(P.u(i, j, k), mul_rn(sub_rn(pc, P.p.ld(i - 1, j, k)), P.g.rdx));
    if (P.g.topo[1] != FLAT) P.v(i, j, k) = sub_rn(P.v(i, j, k), mul_rn(sub_rn(pc, P.p.ld(i, j - 1, k)), P.g.rdy));
    if (P.g.topo[2] != FLAT) P.w(i, j, k) = sub_rn(P.w(i, j, k), mul_rn(sub_rn(pc, P.p.ld(i, j, k - 1)), P.g.rdzF(k)));
}

// K15 + K16 + rescale fused (single-device path of rk3_substep! / ab2_step!): reads the solver output directly,
// p = real(ϕ)·scale (copy_real_component!), u -= ∂x p etc. (make_pressure_correction!), then stores p / Δτ -- the same
// values, in the same order of operations, as the three separate reference kernels; the halo cells of p that the
// correction needs are resolved by index (periodic wrap / no-flux mirror) instead of a halo fill in between.
template <typename T>
struct CorrectFusedP {
    GridD<T> g;
    Fld<T> u, v, w, p;
    const T *sol;
    long ldx, ldxy;
    int cplx, zperm;
    int west;   // the solver output has the west neighbour's column at i = 0 (distributed slab-x): no periodic wrap in x
    T scale, denom;
};
template <typename T>
__global__ void __launch_bounds__(256) correct_fused_kernel(const __grid_constant__ CorrectFusedP<T> P) {
    int i, j, k;
    if (!cell_from_block(P.g.N[0], P.g.N[1], i, j, k)) return;
    auto S = [&](int a, int b, int c) -> T {
        const int cc = P.zperm ? makhoul_index(c - 1, P.g.N[2]) : c - 1;
        const long o = (a - 1) + (b - 1) * P.ldx + (long)cc * P.ldxy;
        return mul_rn(P.cplx ? __ldg(P.sol + 2 * o) : __ldg(P.sol + o), P.scale);   // rounded like the stored p of the reference
    };
    // index of the lower neighbour along d: periodic wrap, or the cell itself where the no-flux halo mirrors it
    auto lower = [&](int idx, int d) { return (idx > 1 || (d == 0 && P.west)) ? idx - 1 : (P.g.topo[d] == PERIODIC ? P.g.N[d] : 1); };
    const T pc = S(i, j, k);
    if (P.g.topo[0] != FLAT) P.u(i, j, k) = sub_rn(P.u(i, j, k), mul_rn(sub_rn(pc, S(lower(i, 0), j, k)), P.g.rdx));
    if (P.g.topo[1] != FLAT) P.v(i, j, k) = sub_rn(P.v(i, j, k), mul_rn(sub_rn(pc, S(i, lower(j, 1), k)), P.g.rdy));
    if (P.g.topo[2] != FLAT) P.w(i, j, k) = sub_rn(P.w(i, j, k), mul_rn(sub_rn(pc, S(i, j, lower(k, 2))), P.g.rdzF(k)));
    P.p(i, j, k) = pc / P.denom;
}

// pNHS ./= Δt over the whole parent array (pressure_correction.jl:101-103)
template <typename T>
__global__ void scale_kernel(T *p, long n, T denom) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = p[t] / denom;
}
template <typename T>
__global__ void fill_kernel(T *p, long n, T v) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}
template <typename T>
__global__ void any_nan_kernel(const T *p, long n, int *flag) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    bool bad = false;
    for (; t < n; t += stride) bad |= isnan(p[t]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

}  // namespace ob
