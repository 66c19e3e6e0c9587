# OceananigansB200Ext.jl -- the reference-side binding of libocean_b200.so (UNTESTED HERE: Julia is not installed in the
# build image or on the GPU box; the same ABI is exercised through the ctypes mirror in oceananigans.jl_b200/_abi.py).
#
# Pattern: exactly that of ext/OceananigansCUDAExt.jl (array / architecture mapping), ext/OceananigansNCCLExt
# (re-dispatching hot-path methods on an architecture alias) and ext/OceananigansReactantExt/TimeSteppers.jl:27-30
# (re-dispatching time_step! on AbstractModel{TS, <:Arch}).  No KernelAbstractions, no CPU fallback on this path.
module OceananigansB200Ext

using Oceananigans
using Oceananigans.Architectures: AbstractSerialArchitecture
using Oceananigans.Grids: RectilinearGrid, topology, Periodic, Bounded, Flat, halo_size
using Oceananigans.Models.NonhydrostaticModels: NonhydrostaticModel
using Oceananigans.TimeSteppers: RungeKutta3TimeStepper, QuasiAdamsBashforth2TimeStepper, tick!, Clock
using Oceananigans.Advection: WENO, Centered
using Oceananigans.TurbulenceClosures: ScalarDiffusivity, Smagorinsky, AnisotropicMinimumDissipation, VerticallyImplicitTimeDiscretization
using Oceananigans.BoundaryConditions: FieldBoundaryConditions, BoundaryCondition, Flux, Value, Gradient, Open, Periodic as PBC
import Oceananigans.Architectures as AC
import Oceananigans.TimeSteppers: time_step!, update_state!, cache_previous_tendencies!
import Oceananigans.BoundaryConditions: fill_halo_regions!
import Oceananigans.Models.NonhydrostaticModels: compute_pressure_correction!, make_pressure_correction!, compute_tendencies!
import Oceananigans.Utils: launch!, sync_device!

const lib = get(ENV, "OCEAN_B200_LIB", "libocean_b200.so")

# ---- status handling: every entry point returns Int32; ob_last_error() gives the message --------------------------------
struct OceanB200Error <: Exception; code::Int32; msg::String; end
@inline function check(status::Int32)
    status == 0 && return nothing
    throw(OceanB200Error(status, unsafe_string(ccall((:ob_last_error, lib), Cstring, ()))))
end
macro ob(f, argtypes, args...)   # @ob ob_sync (Ptr{Cvoid},) ctx
    esc(:(check(ccall(($(QuoteNode(f)), lib), Int32, $argtypes, $(args...)))))
end

# ---- architecture (src/Architectures.jl:21-132) ---------------------------------------------------------------------
mutable struct B200 <: AbstractSerialArchitecture
    device :: Int32
    ctx    :: Ptr{Cvoid}
    function B200(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        @ob ob_init (Int32, Ref{Ptr{Cvoid}}) Int32(device) ref
        arch = new(Int32(device), ref[])
        finalizer(a -> ccall((:ob_shutdown, lib), Int32, (Ptr{Cvoid},), a.ctx), arch)
        return arch
    end
end

# ---- array type: owns device memory obtained from the library (ob_malloc / ob_free) ----------------------------------
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    ptr  :: Ptr{Cvoid}
    dims :: NTuple{N, Int}
    arch :: B200
    function B200Array{T}(arch::B200, dims::NTuple{N, Int}) where {T, N}
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        @ob ob_malloc (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}) arch.ctx prod(dims) * sizeof(T) ref   # zero-initialised
        a = new{T, N}(ref[], dims, arch)
        finalizer(x -> ccall((:ob_free, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), x.arch.ctx, x.ptr), a)  # thread-safe, no NCCL
        return a
    end
end
Base.size(a::B200Array) = a.dims
Base.pointer(a::B200Array) = a.ptr
Base.getindex(::B200Array, I...) = error("scalar indexing of a B200Array is disallowed (cf. allowscalar(false))")
Base.similar(a::B200Array{T}, ::Type{S} = T, dims::Dims = size(a)) where {T, S} = B200Array{S}(a.arch, dims)
function Base.copyto!(dst::B200Array{T}, src::Array{T}) where T
    @ob ob_memcpy_h2d (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) dst.arch.ctx dst.ptr src sizeof(src)
    @ob ob_sync (Ptr{Cvoid},) dst.arch.ctx
    return dst
end
function Base.copyto!(dst::Array{T}, src::B200Array{T}) where T
    @ob ob_memcpy_d2h (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) src.arch.ctx dst src.ptr sizeof(dst)
    return dst
end
Base.copyto!(dst::B200Array{T}, src::B200Array{T}) where T =
    (@ob ob_memcpy_d2d (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) dst.arch.ctx dst.ptr src.ptr prod(size(dst)) * sizeof(T); dst)
Base.Array(a::B200Array{T, N}) where {T, N} = copyto!(Array{T, N}(undef, size(a)), a)
Base.fill!(a::B200Array{T}, v) where T =
    (@ob ob_fill (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Float64) a.arch.ctx a.ptr length(a) ftype(T) Float64(v); a)
function Base.any(::typeof(isnan), a::B200Array{T}) where T       # Diagnostics/nan_checker.jl
    flag = Ref{Int32}(0)
    @ob ob_any_nan (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Ref{Int32}) a.arch.ctx a.ptr length(a) ftype(T) flag
    return flag[] != 0
end
ftype(::Type{Float32}) = Int32(0)
ftype(::Type{Float64}) = Int32(1)

AC.device(a::B200) = a.device
AC.array_type(::B200) = B200Array
AC.architecture(a::B200Array) = a.arch
AC.on_architecture(arch::B200, a::Array{T}) where T = copyto!(B200Array{T}(arch, size(a)), a)
AC.on_architecture(::AC.CPU, a::B200Array) = Array(a)
AC.on_architecture(::B200, a::B200Array) = a
AC.synchronize(a::B200) = @ob ob_sync (Ptr{Cvoid},) a.ctx
sync_device!(a::B200) = AC.synchronize(a)
Base.zeros(arch::B200, FT, N...) = B200Array{FT}(arch, Tuple(Int.(N)))       # Grids/zeros_and_ones.jl:8
Oceananigans.Advection.default_weno_weight_computation(::B200) = Oceananigans.Utils.BackendOptimizedDivision
# nothing may silently fall back to KernelAbstractions on this architecture:
launch!(::B200, args...; kw...) = error("no KernelAbstractions path on B200: this operation is outside the accelerated hot path")

# ---- POD descriptors: field-for-field mirrors of include/ocean_b200.h ------------------------------------------------
struct ObGridDesc
    float_type::Int32; N::NTuple{3, Int32}; H::NTuple{3, Int32}; topology::NTuple{3, Int32}
    L::NTuple{3, Float64}; d::NTuple{3, Float64}
    dzf_host::Ptr{Cvoid}; dzc_host::Ptr{Cvoid}; n_dzf::Int32; n_dzc::Int32
end
struct ObBcDesc; kind::NTuple{6, Int32}; value::NTuple{6, Float64}; end
struct ObClosureDesc
    kind::Int32; nu::Float64; kappa::NTuple{8, Float64}; cs::Float64; lilly::Int32; cb::Float64; Pr::NTuple{8, Float64}
    Cnu::Float64; Ckappa::NTuple{8, Float64}; amd_has_cb::Int32; vertically_implicit::Int32
end
struct ObModelDesc
    grid::ObGridDesc
    advection_kind::Int32; advection_order::Int32; weno_division::Int32
    n_closures::Int32; closures::NTuple{4, ObClosureDesc}
    buoyancy_kind::Int32; buoyancy_tracer::Int32; temperature_tracer::Int32; salinity_tracer::Int32
    g::Float64; thermal_expansion::Float64; haline_contraction::Float64
    has_coriolis::Int32; f::Float64
    n_tracers::Int32; stepper::Int32; chi::Float64; has_hydrostatic_pressure::Int32
    bcs_u::ObBcDesc; bcs_v::ObBcDesc; bcs_w::ObBcDesc; bcs_p::ObBcDesc; bcs_phy::ObBcDesc
    bcs_tracer::NTuple{8, ObBcDesc}; bcs_nue::NTuple{4, ObBcDesc}; bcs_kappae::NTuple{4, NTuple{8, ObBcDesc}}
end

topo_id(::Type{Periodic}) = Int32(0); topo_id(::Type{Bounded}) = Int32(1); topo_id(::Type{Flat}) = Int32(2)

"Grid descriptor: spacings are passed exactly as Julia constructed them (grid_generation.jl:34-156)."
function grid_desc(grid::RectilinearGrid{FT}) where FT
    TX, TY, TZ = topology(grid)
    stretched = !(grid.z.Δᵃᵃᶜ isa Number)
    (grid.Δxᶠᵃᵃ isa Number && grid.Δyᵃᶠᵃ isa Number) || throw(ArgumentError("B200: only z may be variably spaced"))
    dzf = stretched ? Array(parent(grid.z.Δᵃᵃᶠ)) : FT[]
    dzc = stretched ? Array(parent(grid.z.Δᵃᵃᶜ)) : FT[]
    desc = ObGridDesc(ftype(FT), Int32.((grid.Nx, grid.Ny, grid.Nz)), Int32.((grid.Hx, grid.Hy, grid.Hz)),
                      (topo_id(TX), topo_id(TY), topo_id(TZ)), Float64.((grid.Lx, grid.Ly, grid.Lz)),
                      Float64.((grid.Δxᶠᵃᵃ, grid.Δyᵃᶠᵃ, stretched ? 0 : grid.z.Δᵃᵃᶠ)),
                      stretched ? pointer(dzf) : C_NULL, stretched ? pointer(dzc) : C_NULL, Int32(length(dzf)), Int32(length(dzc)))
    return desc, (dzf, dzc)   # keep the host arrays alive across ob_model_create
end

# ---- translation of the model description; anything the ABI cannot express throws (no CPU fallback) -----------------
unsupported(what) = throw(ArgumentError("B200: $what is outside the accelerated NonhydrostaticModel path (SURVEY.md §2)"))

bc_kind(::Nothing) = (Int32(0), 0.0)
function bc_kind(bc::BoundaryCondition)
    c = bc.classification
    c isa PBC && return (Int32(1), 0.0)
    v = bc.condition
    # arrays (and time-independent boundary functions tabulated by the user into arrays) go through ob_model_set_bc_array
    (v isa Number || v === nothing || v isa AbstractArray) || unsupported("a function-valued boundary condition (tabulate it into an array)")
    val = (v === nothing || v isa AbstractArray) ? 0.0 : Float64(v)
    c isa Flux     && return (Int32(2), val)
    c isa Value    && return (Int32(3), val)
    c isa Gradient && return (Int32(4), val)
    c isa Open     && (v === nothing ? (return (Int32(5), 0.0)) : unsupported("an open boundary with a prescribed value"))
    unsupported("boundary condition $(typeof(c))")
end
function bc_desc(bcs::FieldBoundaryConditions)
    ks = map(bc_kind, (bcs.west, bcs.east, bcs.south, bcs.north, bcs.bottom, bcs.top))
    return ObBcDesc(ntuple(i -> ks[i][1], 6), ntuple(i -> ks[i][2], 6))
end
bc_desc(::Nothing) = ObBcDesc(ntuple(_ -> Int32(0), 6), ntuple(_ -> 0.0, 6))

pad8(t) = ntuple(i -> i <= length(t) ? Float64(t[i]) : 0.0, 8)
closure_desc(c::ScalarDiffusivity, names) =
    (c.ν isa Number && all(κ -> κ isa Number, values(c.κ))) ?
        ObClosureDesc(1, Float64(c.ν), pad8(values(c.κ)), 0.0, 0, 0.0, pad8(()), 0.0, pad8(()), 0,
                      Int32(Oceananigans.TimeSteppers.time_discretization(c) isa VerticallyImplicitTimeDiscretization)) : unsupported("a function-valued diffusivity")
function closure_desc(c::Smagorinsky, names)
    coeff = c.coefficient
    lilly = !(coeff isa Number)
    cs = lilly ? coeff.smagorinsky : coeff
    cb = lilly ? coeff.reduction_factor : 0.0
    (cs isa Number) || unsupported("DynamicSmagorinsky")
    return ObClosureDesc(2, 0.0, pad8(()), Float64(cs), Int32(lilly), Float64(cb), pad8(values(c.Pr)), 0.0, pad8(()), 0, 0)
end
closure_desc(c::AnisotropicMinimumDissipation, names) =
    ObClosureDesc(3, 0.0, pad8(()), 0.0, 0, c.Cb === nothing ? 0.0 : Float64(c.Cb), pad8(()), Float64(c.Cν), pad8(values(c.Cκ)), Int32(c.Cb !== nothing), 0)
closure_desc(c, names) = unsupported("closure $(typeof(c))")

function model_desc(model::NonhydrostaticModel)
    grid = model.grid
    gd, keep = grid_desc(grid)
    adv = model.advection.momentum   # the shim requires one scheme for momentum and tracers
    kind, order = adv isa WENO ? (Int32(2), 2 * Oceananigans.Advection.required_halo_size_x(adv) - 1) :
                  adv isa Centered ? (Int32(1), 2 * Oceananigans.Advection.required_halo_size_x(adv)) : unsupported("advection $(typeof(adv))")
    closures = model.closure === nothing ? () : model.closure isa Tuple ? model.closure : (model.closure,)
    length(closures) <= 4 || unsupported("more than 4 closures")
    names = keys(model.tracers)
    length(names) <= 8 || unsupported("more than 8 tracers")
    cds = ntuple(i -> i <= length(closures) ? closure_desc(closures[i], names) : ObClosureDesc(0, 0.0, pad8(()), 0.0, 0, 0.0, pad8(()), 0.0, pad8(()), 0), 4)
    b = model.buoyancy === nothing ? nothing : model.buoyancy.formulation
    bk, ib, iT, iS, g, α, β = Int32(0), Int32(0), Int32(0), Int32(0), 0.0, 0.0, 0.0
    if b isa Oceananigans.BuoyancyFormulations.BuoyancyTracer
        bk, ib = Int32(1), Int32(findfirst(==(:b), names) - 1)
    elseif b isa Oceananigans.BuoyancyFormulations.SeawaterBuoyancy
        eos = b.equation_of_state
        eos isa Oceananigans.BuoyancyFormulations.LinearEquationOfState || unsupported("a nonlinear equation of state")
        bk, iT, iS = Int32(2), Int32(findfirst(==(:T), names) - 1), Int32(findfirst(==(:S), names) - 1)
        g, α, β = Float64(b.gravitational_acceleration), Float64(eos.thermal_expansion), Float64(eos.haline_contraction)
    elseif b !== nothing
        unsupported("buoyancy $(typeof(b))")
    end
    cor = model.coriolis
    (cor === nothing || cor isa Oceananigans.Coriolis.FPlane) || unsupported("Coriolis $(typeof(cor))")
    all(f -> f isa Oceananigans.Forcings.zeroforcing |> typeof || f === Oceananigans.Forcings.zeroforcing, values(model.forcing)) || unsupported("user forcing functions")
    ts = model.timestepper
    tr_bcs = ntuple(i -> i <= length(names) ? bc_desc(model.tracers[i].boundary_conditions) : bc_desc(nothing), 8)
    K = model.closure_fields
    nue_bcs = ntuple(i -> bc_desc(nothing), 4); kap_bcs = ntuple(i -> ntuple(j -> bc_desc(nothing), 8), 4)   # defaults are filled in by the library
    pHY = model.pressures.pHY′
    desc = ObModelDesc(gd, kind, Int32(order), Int32(1), Int32(length(closures)), cds, bk, ib, iT, iS, g, α, β,
                       Int32(cor !== nothing), cor === nothing ? 0.0 : Float64(cor.f), Int32(length(names)),
                       ts isa RungeKutta3TimeStepper ? Int32(0) : Int32(1), ts isa RungeKutta3TimeStepper ? 0.1 : Float64(ts.χ),
                       Int32(pHY !== nothing), bc_desc(model.velocities.u.boundary_conditions), bc_desc(model.velocities.v.boundary_conditions),
                       bc_desc(model.velocities.w.boundary_conditions), bc_desc(model.pressures.pNHS.boundary_conditions),
                       pHY === nothing ? bc_desc(nothing) : bc_desc(pHY.boundary_conditions), tr_bcs, nue_bcs, kap_bcs)
    return desc, keep
end

function create_handle(model)
    desc, keep = model_desc(model)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep @ob ob_model_create (Ptr{Cvoid}, Ref{ObModelDesc}, Ref{Ptr{Cvoid}}) model.architecture.ctx Ref(desc) ref
    h = ref[]
    bind!(h, 0, model.velocities.u); bind!(h, 1, model.velocities.v); bind!(h, 2, model.velocities.w)
    bind!(h, 3, model.pressures.pNHS)
    model.pressures.pHY′ === nothing || bind!(h, 4, model.pressures.pHY′)
    for (t, c) in enumerate(model.tracers); bind!(h, 16 + t - 1, c); end
    Gⁿ, G⁻ = model.timestepper.Gⁿ, model.timestepper.G⁻
    for (n, (a, b)) in enumerate(zip(Gⁿ, G⁻)); bind!(h, 32 + n - 1, a); bind!(h, 48 + n - 1, b); end
    closures = model.closure === nothing ? () : model.closure isa Tuple ? model.closure : (model.closure,)
    for (m, c) in enumerate(closures)
        K = model.closure isa Tuple ? model.closure_fields[m] : model.closure_fields
        c isa ScalarDiffusivity && continue
        bind!(h, 64 + m - 1, K.νₑ)
        if c isa AnisotropicMinimumDissipation
            for (t, κ) in enumerate(K.κₑ); bind!(h, 80 + (m - 1) * 8 + t - 1, κ); end
        end
    end
    # array-valued conditions: condition[i, j] over the boundary plane, already a device array (on_architecture(::B200, ...))
    for (id, f) in ((0, model.velocities.u), (1, model.velocities.v), (2, model.velocities.w), ((16 + t - 1, c) for (t, c) in enumerate(model.tracers))...)
        bcs = f.boundary_conditions
        for (side, bc) in enumerate((bcs.west, bcs.east, bcs.south, bcs.north, bcs.bottom, bcs.top))
            (bc isa BoundaryCondition && bc.condition isa AbstractArray) || continue
            @ob ob_model_set_bc_array (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}) h Int32(id) Int32(side - 1) pointer(bc.condition)
        end
    end
    finalizer(_ -> ccall((:ob_model_destroy, lib), Int32, (Ptr{Cvoid},), h), model.timestepper)
    return h
end

# ---- model handle cached on the Julia model --------------------------------------------------------------------------
const HANDLES = IdDict{Any, Ptr{Cvoid}}()
const B200Model{TS} = NonhydrostaticModel{TS, <:Any, <:B200}
handle(model) = get!(() -> create_handle(model), HANDLES, model)

function bind!(h, id::Integer, f)   # ob_model_bind_field(model, field_id, device pointer of parent(f))
    @ob ob_model_bind_field (Ptr{Cvoid}, Int32, Ptr{Cvoid}) h Int32(id) pointer(parent(f))
end

# ---- hot-path method overrides (SURVEY.md §8b item 3) ----------------------------------------------------------------
function time_step!(model::B200Model{<:RungeKutta3TimeStepper}, Δt; callbacks = [])
    first = model.clock.iteration == 0
    if isempty(callbacks)
        @ob ob_time_step_rk3 (Ptr{Cvoid}, Float64, Int32) handle(model) Float64(Δt) Int32(first)     # one call per step
    else                                                                                            # host-driven stages
        ts = model.timestepper
        first && update_state!(model, callbacks)
        for (γ, ζ) in ((ts.γ¹, nothing), (ts.γ², ts.ζ²), (ts.γ³, ts.ζ³))
            @ob ob_rk3_substep (Ptr{Cvoid}, Float64, Float64, Float64, Int32) handle(model) Float64(Δt) Float64(γ) Float64(something(ζ, 0)) Int32(!isnothing(ζ))
            cache_previous_tendencies!(model)
            update_state!(model, callbacks)
        end
    end
    tick!(model.clock, Δt)
    return nothing
end

function time_step!(model::B200Model{<:QuasiAdamsBashforth2TimeStepper}, Δt; callbacks = [], euler = false)
    euler = euler | (Δt != model.clock.last_Δt)
    @ob ob_time_step_ab2 (Ptr{Cvoid}, Float64, Int32, Int32) handle(model) Float64(Δt) Int32(euler) Int32(model.clock.iteration == 0)
    tick!(model.clock, Δt)
    return nothing
end

update_state!(model::B200Model, callbacks = []; kw...) = (@ob ob_update_state (Ptr{Cvoid},) handle(model); foreach(c -> c(model), callbacks))
compute_tendencies!(model::B200Model, callbacks = []) = @ob ob_compute_tendencies (Ptr{Cvoid},) handle(model)
cache_previous_tendencies!(model::B200Model) = @ob ob_cache_tendencies (Ptr{Cvoid},) handle(model)
compute_pressure_correction!(model::B200Model, Δt) = @ob ob_compute_pressure_correction (Ptr{Cvoid}, Float64) handle(model) Float64(Δt)
make_pressure_correction!(model::B200Model, Δt) = @ob ob_make_pressure_correction (Ptr{Cvoid}, Float64) handle(model) Float64(Δt)

# fill_halo_regions! of a prognostic / pressure field of a B200 model (fill_halo_regions.jl:20-38)
function fill_model_halo!(model::B200Model, field_id::Integer; fill_normal_flow_bcs = true)
    @ob ob_fill_halo (Ptr{Cvoid}, Int32, Int32) handle(model) Int32(field_id) Int32(fill_normal_flow_bcs)
end

# TimeStepWizard (cell_advection_timescale.jl:14-35): device min-reduction
function Oceananigans.Advection.cell_advection_timescale(model::B200Model)
    τ = Ref{Float64}(0)
    @ob ob_cell_advection_timescale (Ptr{Cvoid}, Ref{Float64}) handle(model) τ
    return τ[]
end

end # module
