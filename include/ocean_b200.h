/*
 * ocean_b200.h -- C ABI of libocean_b200.so: the B200-native (sm_100a) NonhydrostaticModel time step
 * of Oceananigans.jl on a RectilinearGrid.
 *
 * This is the drop-in boundary.  Oceananigans has no plugin registry: the plugin API is
 * Julia multiple dispatch on the architecture type parameter of the grid / model
 * (src/Architectures.jl:21-132, src/Oceananigans.jl:190, nonhydrostatic_model.jl:32-33).  A thin Julia
 * extension (oceananigans.jl_b200/julia/OceananigansB200Ext.jl) adds a `B200` architecture and re-dispatches
 * the hot-path methods onto the entry points below with `ccall`; INTEGRATION.md shows every binding.
 *
 * Conventions
 *   - every function returns int32 status: 0 = OB_OK, negative = error; ob_last_error() returns the message.
 *   - opaque handles; plain pointers and sizes only; no callbacks into the host language.
 *   - the HOST owns field memory: it is allocated with ob_malloc (device memory) and bound to a model with
 *     ob_model_bind_field.  Arrays have the reference layout: one contiguous parent array, x fastest, of
 *     size (Nx+2Hx, Ny+2Hy, Nz+2Hz), +1 along a Bounded direction for Face-located fields, and size N
 *     (no halo) along Flat directions (src/Grids/new_data.jl:11-74).
 *   - the library owns its workspaces (spectral storage, cuFFT plans, streams, communicators).
 *   - all calls are asynchronous on the context's stream except ob_memcpy_d2h, ob_any_nan, ob_sync,
 *     ob_reduce_* .
 *   - there is NO CPU fallback: every entry point fails with OB_ERR_NO_DEVICE when no sm_100 device exists.
 */
#ifndef OCEAN_B200_H
#define OCEAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OB_OK 0
#define OB_ERR_INVALID (-1)
#define OB_ERR_CUDA (-2)
#define OB_ERR_UNSUPPORTED (-3) /* option outside the hot-path scope: the shim must raise */
#define OB_ERR_NO_DEVICE (-4)
#define OB_ERR_UNBOUND (-5)
#define OB_ERR_CUFFT (-6)
#define OB_ERR_NCCL (-7)

#define OB_MAX_TRACERS 8
#define OB_MAX_CLOSURES 4

typedef struct ob_ctx ob_ctx;
typedef struct ob_model ob_model;
typedef struct ob_solver ob_solver;

typedef enum { OB_F32 = 0, OB_F64 = 1 } ob_float_type;
/* src/Grids: Periodic / Bounded / Flat */
typedef enum { OB_PERIODIC = 0, OB_BOUNDED = 1, OB_FLAT = 2 } ob_topology;
/* src/BoundaryConditions/boundary_condition_classifications.jl */
typedef enum {
    OB_BC_NONE = 0,       /* `nothing` (Flat, or auxiliary Face field on a Bounded side) */
    OB_BC_PERIODIC = 1,
    OB_BC_FLUX = 2,       /* Flux; value 0 == NoFlux (the Bounded+Center default) */
    OB_BC_VALUE = 3,
    OB_BC_GRADIENT = 4,
    OB_BC_IMPENETRABLE = 5, /* Open(nothing): the Bounded+Face default */
    OB_BC_COMMUNICATION = 6 /* distributed halo: filled by ob_dist halo exchange */
} ob_bc_kind;
typedef enum { OB_ADV_NONE = 0, OB_ADV_CENTERED = 1, OB_ADV_WENO = 2 } ob_advection_kind;
typedef enum { OB_CLOSURE_SCALAR_DIFFUSIVITY = 1, OB_CLOSURE_SMAGORINSKY = 2, OB_CLOSURE_AMD = 3 } ob_closure_kind;
typedef enum { OB_BUOYANCY_NONE = 0, OB_BUOYANCY_TRACER = 1, OB_BUOYANCY_LINEAR_SEAWATER = 2 } ob_buoyancy_kind;
typedef enum { OB_RK3 = 0, OB_AB2 = 1 } ob_stepper_kind;
/* src/Utils/newton_div.jl:34-57 ; ext/OceananigansCUDAExt.jl:147-171 */
typedef enum { OB_DIV_EXACT = 0, OB_DIV_RCP_NEWTON = 1 } ob_weno_division;

/* Field ids for ob_model_bind_field / ob_fill_halo.  Tracer t: OB_FIELD_TRACER0 + t, etc. */
enum {
    OB_FIELD_U = 0, OB_FIELD_V = 1, OB_FIELD_W = 2,
    OB_FIELD_PNHS = 3, OB_FIELD_PHY = 4,
    OB_FIELD_TRACER0 = 16,               /* +t                                   */
    OB_FIELD_GN0 = 32,                   /* +n, n = 0..2 velocities, 3+t tracers */
    OB_FIELD_GM0 = 48,                   /* previous tendencies G⁻               */
    OB_FIELD_NUE0 = 64,                  /* +closure index                       */
    OB_FIELD_KAPPAE0 = 80                /* + closure*OB_MAX_TRACERS + t         */
};

/* RectilinearGrid (src/Grids/rectilinear_grid.jl:3-24).  Spacing arrays are passed exactly as the host
 * constructed them (grid_generation.jl:34-156): for a stretched z, dzf has Nz+2Hz+1 entries with logical
 * index k at dzf[k+Hz]; dzc has Nz+2Hz entries (Bounded z; one fewer when z is Periodic) with logical k at
 * dzc[k+Hz-1].  NULL => regular. */
typedef struct {
    int32_t float_type;     /* ob_float_type */
    int32_t N[3], H[3];
    int32_t topology[3];    /* ob_topology */
    double L[3];            /* extents Lx, Ly, Lz (as rounded to the grid float type) */
    double d[3];            /* regular spacings Δx, Δy, Δz (as rounded to the grid float type) */
    const void *dzf_host;   /* optional stretched-z spacings (host pointers, grid float type) */
    const void *dzc_host;
    int32_t n_dzf, n_dzc;
} ob_grid_desc;

typedef struct {
    int32_t kind[6];        /* west, east, south, north, bottom, top : ob_bc_kind */
    double value[6];        /* constant flux / value / gradient */
} ob_bc_desc;

typedef struct {
    int32_t kind;           /* ob_closure_kind */
    double nu;              /* ScalarDiffusivity ν */
    double kappa[OB_MAX_TRACERS];
    double cs;              /* Smagorinsky coefficient */
    int32_t lilly;          /* SmagorinskyLilly */
    double cb;              /* Lilly reduction factor / AMD Cb */
    double Pr[OB_MAX_TRACERS];
    double Cnu;             /* AMD Poincaré constants */
    double Ckappa[OB_MAX_TRACERS];
    int32_t amd_has_cb;
    int32_t vertically_implicit; /* ScalarDiffusivity(VerticallyImplicitTimeDiscretization(); ...) (scalar_diffusivity.jl:113-137):
                                  * z-Bounded grids only; the substeps then run implicit_step! on every prognostic field */
    /* Smagorinsky(coefficient = DynamicCoefficient(averaging = dims; minimum_numerator)) -- DynamicSmagorinsky with a
     * directionally averaged coefficient (dynamic_coefficient.jl:107-118, 208-212, 306-351).  dynamic != 0: `cs` is ignored and
     * c_s^2 = max(<LM>, minimum_numerator) / <MM> is recomputed in every update_state! (schedule IterationInterval(1));
     * averaging_dims: bit d-1 set = dimension d is averaged ((1, 2) -> 3, Colon / (1, 2, 3) -> 7).  LagrangianAveraging,
     * other schedules, Flat directions and distributed grids: OB_ERR_UNSUPPORTED. */
    int32_t dynamic;
    int32_t averaging_dims;
    double minimum_numerator;
} ob_closure_desc;

/* NonhydrostaticModel(grid; advection, closure, buoyancy, coriolis, tracers, timestepper)
 * (src/Models/NonhydrostaticModels/nonhydrostatic_model.jl:124-313) */
typedef struct {
    ob_grid_desc grid;
    int32_t advection_kind;   /* ob_advection_kind */
    int32_t advection_order;  /* WENO: 3,5,7,9,11 ; Centered: 2,4,...,12 -- the buffers 1..6 of src/Advection/Advection.jl:52 (anything else: OB_ERR_UNSUPPORTED) */
    int32_t weno_division;    /* ob_weno_division */
    int32_t n_closures;
    ob_closure_desc closures[OB_MAX_CLOSURES];
    int32_t buoyancy_kind;    /* ob_buoyancy_kind */
    int32_t buoyancy_tracer;  /* index of b        */
    int32_t temperature_tracer, salinity_tracer;
    double g, thermal_expansion, haline_contraction;
    int32_t has_coriolis;
    double f;                 /* FPlane f */
    int32_t n_tracers;
    int32_t stepper;          /* ob_stepper_kind */
    double chi;               /* AB2 χ */
    int32_t has_hydrostatic_pressure; /* pHY′ exists (buoyancy and z not Periodic) */
    ob_bc_desc bcs_u, bcs_v, bcs_w, bcs_p, bcs_phy;
    ob_bc_desc bcs_tracer[OB_MAX_TRACERS];
    ob_bc_desc bcs_nue[OB_MAX_CLOSURES];
    ob_bc_desc bcs_kappae[OB_MAX_CLOSURES][OB_MAX_TRACERS];
} ob_model_desc;

/* ---- context & memory (replaces CUDA.jl array plumbing: ext/OceananigansCUDAExt.jl:46-82) ------------- */
int32_t ob_init(int32_t device, ob_ctx **ctx);
int32_t ob_shutdown(ob_ctx *ctx);
const char *ob_last_error(void);
int32_t ob_device_count(int32_t *n);
int32_t ob_sync(ob_ctx *ctx);                                         /* Architectures.synchronize / sync_device! */
/* CUDA-event timer on the context's stream: start records, stop records + synchronises and returns milliseconds */
int32_t ob_timer_start(ob_ctx *ctx);
int32_t ob_timer_stop(ob_ctx *ctx, double *ms);
/* measured FP64 FMA issue rate of the device (thread-level instructions per second): the second roofline of the
 * FP64-bound tendency kernel in bench.py */
int32_t ob_fp64_peak(ob_ctx *ctx, double *instr_per_s);
int32_t ob_malloc(ob_ctx *ctx, size_t bytes, void **ptr);             /* Base.zeros(::B200, FT, dims...) */
int32_t ob_free(ob_ctx *ctx, void *ptr);                              /* finalizer / unsafe_free! */
int32_t ob_malloc_host(ob_ctx *ctx, size_t bytes, void **ptr);        /* pinned staging buffers */
int32_t ob_free_host(ob_ctx *ctx, void *ptr);
int32_t ob_memcpy_h2d(ob_ctx *ctx, void *dst, const void *src, size_t bytes); /* on_architecture(::B200, ::Array) */
int32_t ob_memcpy_d2h(ob_ctx *ctx, void *dst, const void *src, size_t bytes); /* on_architecture(::CPU, ::B200Array) */
int32_t ob_memcpy_d2d(ob_ctx *ctx, void *dst, const void *src, size_t bytes); /* copyto! / device_copy_to! */
/* stream-ordered copy into PINNED host memory; the host may read dst after ob_sync(ctx) */
int32_t ob_memcpy_d2h_async(ob_ctx *ctx, void *dst, const void *src, size_t bytes);
/* later calls on ctx wait (on the device) for everything submitted so far on `other` (same device); the host is not
 * blocked.  Joins the lanes of a host-streamed ensemble: several contexts = several streams whose copies and kernels
 * overlap (no reference counterpart: the reference's CPU() path has no host<->device boundary) */
int32_t ob_stream_wait(ob_ctx *ctx, ob_ctx *other);
int32_t ob_fill(ob_ctx *ctx, void *ptr, size_t n, int32_t float_type, double value); /* fill! */
int32_t ob_any_nan(ob_ctx *ctx, const void *ptr, size_t n, int32_t float_type, int32_t *flag); /* NaNChecker */
/* Diagnostics used by TimeStepWizard (src/Advection/cell_advection_timescale.jl:14-35) */
int32_t ob_cell_advection_timescale(ob_model *m, double *tau);

/* ---- model --------------------------------------------------------------------------------------------- */
int32_t ob_model_create(ob_ctx *ctx, const ob_model_desc *desc, ob_model **model);
int32_t ob_model_destroy(ob_model *m);
int32_t ob_model_bind_field(ob_model *m, int32_t field_id, void *device_ptr);
/* Array-valued boundary condition (BoundaryCondition with an AbstractArray condition: getbc(bc, i, j, ...) = condition[i, j],
 * src/BoundaryConditions/boundary_condition.jl): `values` is a device array (ob_malloc) of the field's interior extent in
 * the two tangential dimensions of `side` (0 west .. 5 top), the lower dimension fastest, in the grid float type; it
 * replaces the constant of the Flux / Value / Gradient condition declared for that side.  NULL restores the constant.
 * The caller keeps ownership.  This is also how the Julia shim passes time-independent boundary FUNCTIONS: tabulated once
 * on the host. */
int32_t ob_model_set_bc_array(ob_model *m, int32_t field_id, int32_t side, const void *values);
/* fill_halo_regions!(field) -- src/BoundaryConditions/fill_halo_regions.jl:20-38 */
int32_t ob_fill_halo(ob_model *m, int32_t field_id, int32_t fill_normal_flow_bcs);
/* fill_halo_regions!(c::OffsetArray, bcs, indices, loc, grid) for ANY field array, model-bound or not (diagnostics, user
 * fields, closure fields built by the host): the same kernels and the same ordering (non-periodic sides first, then
 * periodic over the full parent extent) as ob_fill_halo.  `loc[d]` = 1 for a Face location along d; `bc_arrays` is NULL or six
 * device pointers (NULL entries = constant conditions) laid out as for ob_model_set_bc_array.  Communication halos of a
 * distributed field belong to a model (ob_fill_halo).  src/BoundaryConditions/fill_halo_regions.jl:20-38 */
int32_t ob_fill_halo_array(ob_ctx *ctx, const ob_grid_desc *grid, void *device_ptr, const int32_t *loc, const ob_bc_desc *bcs,
                           const void *const *bc_arrays, int32_t fill_normal_flow_bcs);
/* update_state!(model) -- update_nonhydrostatic_model_state.jl:22-62 (halos, closure fields, pHY′, tendencies) */
int32_t ob_update_state(ob_model *m);
/* compute_tendencies!(model) -- compute_nonhydrostatic_tendencies.jl:12-40 */
int32_t ob_compute_tendencies(ob_model *m);
/* compute_closure_fields! / update_hydrostatic_pressure! */
int32_t ob_compute_closure_fields(ob_model *m);
int32_t ob_update_hydrostatic_pressure(ob_model *m);
/* rk3_substep!(model, Δt, γ, ζ) -- nonhydrostatic_rk3_substep.jl:31-63 ; has_zeta = 0 for the first stage */
int32_t ob_rk3_substep(ob_model *m, double dt, double gamma, double zeta, int32_t has_zeta);
/* ab2_step!(model, Δt) with the model's current χ -- nonhydrostatic_ab2_step.jl:10-57 */
int32_t ob_ab2_step(ob_model *m, double dt, double chi);
/* cache_previous_tendencies! -- cache_nonhydrostatic_tendencies.jl:21-31 */
int32_t ob_cache_tendencies(ob_model *m);
/* compute_pressure_correction! + make_pressure_correction! -- pressure_correction.jl:6-106 */
int32_t ob_compute_pressure_correction(ob_model *m, double dtau);
int32_t ob_make_pressure_correction(ob_model *m, double dtau);
/* whole steps: time_step!(model, Δt) -- runge_kutta_3.jl:103-168 / quasi_adams_bashforth_2.jl:90-126.
 * `first` != 0 replays maybe_prepare_first_time_step! (an extra update_state!). */
int32_t ob_time_step_rk3(ob_model *m, double dt, int32_t first);
int32_t ob_time_step_ab2(ob_model *m, double dt, int32_t euler, int32_t first);
/* implementation options (testing / profiling): OB_OPT_TENDENCY_KERNEL = 0 auto (the staged-ring kernel where it applies,
 * else the marching kernel), 1 generic one-thread-per-cell kernel, 2 flux-sharing marching kernel, 3 marching kernel with
 * TMA-staged stencil planes, 8 staged-ring kernel (tendency_stage.cuh; falls back to 0 where it does not apply) */
#define OB_OPT_TENDENCY_KERNEL 1
/* OB_OPT_FUSE_PROJECTION = 1 (default): single-device substeps fuse real-copy + correction + p rescale; 0: reference kernel sequence */
#define OB_OPT_FUSE_PROJECTION 2
/* OB_OPT_OVERLAP_HALO = 1: distributed update_state! computes the tendency tiles that read no x halo while the halo
 * slabs are in flight (interleave_communication_and_computation.jl:36-74); 0 (default): exchange first, then one launch
 * -- measured faster at 256^3 per GPU on NVLink, where an exchange costs ~50 us and the split launch ~70 us */
#define OB_OPT_OVERLAP_HALO 3
/* OB_OPT_VECTOR_STREAMS = 1 (default): the update, Poisson-source and fused-projection kernels use 128-bit accesses
 * (csrc/streaming.cuh), triply periodic halos are filled in one launch, and ob_time_step_rk3 swaps the roles of the two
 * tendency sets between stages instead of copying; 0: the one-cell-per-thread kernels, the z / y / x fill sequence and the
 * copies of the reference (bit-identical results; kept for A/B checks) */
#define OB_OPT_VECTOR_STREAMS 4
int32_t ob_model_set_option(ob_model *m, int32_t option, int32_t value);
/* number of kernels/library launches issued by this model so far (bench.py's gpu_launches) */
int32_t ob_launch_count(ob_model *m, int64_t *n);
/* per-phase device timing (CUDA events on the model's stream); phase names via ob_phase_name */
int32_t ob_enable_timing(ob_model *m, int32_t enable);
int32_t ob_phase_count(int32_t *n);
const char *ob_phase_name(int32_t phase);
int32_t ob_phase_time_ms(ob_model *m, int32_t phase, double *ms, int64_t *calls);
int32_t ob_reset_timing(ob_model *m);

/* ---- pressure solvers (src/Solvers) --------------------------------------------------------------------- */
/* nonhydrostatic_pressure_solver(arch, grid): FFTBasedPoissonSolver for regular grids,
 * FourierTridiagonalPoissonSolver for stretched z (NonhydrostaticModels.jl:30-52) */
int32_t ob_solver_create(ob_ctx *ctx, const ob_grid_desc *grid, ob_solver **solver);
int32_t ob_solver_destroy(ob_solver *s);
/* solve!(ϕ, solver, rhs): rhs = real (Nx,Ny,Nz) array without halos (already × Δzᶜ for the tridiagonal
 * solver); phi = real (Nx,Ny,Nz) array without halos.  fft_based_poisson_solver.jl:94-124 */
int32_t ob_poisson_solve(ob_solver *s, const void *rhs, void *phi);
/* BatchedTridiagonalSolver (batched_tridiagonal_solver.jl:211-243): z-direction, real or complex rhs */
int32_t ob_batched_tridiagonal_solve(ob_ctx *ctx, int32_t float_type, int32_t is_complex, int32_t Nx, int32_t Ny,
                                     int32_t Nz, const void *a, const void *b, const void *c, const void *f,
                                     void *phi, void *scratch);

/* ---- single-node multi-GPU: slab-x Distributed(B200(); partition=Partition(R)) --------------------------- */
/* One process per GPU.  `nccl_unique_id` (128 bytes) is created by rank 0 with ob_dist_unique_id and
 * broadcast by the host's own plumbing (MPI.Bcast in Julia, torch.distributed in the Python mirror),
 * exactly like ext/OceananigansNCCLExt/nccl_communicator.jl:25-63. */
int32_t ob_dist_unique_id(void *nccl_unique_id_128);
int32_t ob_dist_init(ob_ctx *ctx, int32_t rank, int32_t world, const void *nccl_unique_id_128);
int32_t ob_dist_finalize(ob_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* OCEAN_B200_H */
