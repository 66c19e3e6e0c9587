# OceananigansB200Ext.jl -- the reference-side binding of libocean_b200.so.
#
# NOT EXECUTED HERE: Julia is not installed in the build image or on the GPU box.  The same ABI is exercised through the ctypes
# mirror (oceananigans.jl_b200/_abi.py), and tests/test_abi.py checks the struct field lists and constructor arities of THIS
# file against include/ocean_b200.h.  INTEGRATION.md walks NonhydrostaticModel(grid; ...), set!, time_step! line by line
# and names the method of this file each reference call lands on.
#
# Pattern: that of ext/OceananigansCUDAExt.jl (array / architecture mapping; `const CUDAGPU = GPU{<:CUDABackend}`),
# ext/OceananigansAMDGPUExt.jl (`const ROCGPU = GPU{ROCBackend}`), ext/OceananigansNCCLExt/nccl_distributed.jl:60-134
# (re-dispatching hot-path methods on an architecture alias) and ext/OceananigansReactantExt/TimeSteppers.jl:27-30
# (re-dispatching time_step! on AbstractModel{TS, <:Arch}).  No KernelAbstractions kernel runs on this architecture.
module OceananigansB200Ext

using Oceananigans
using OffsetArrays: OffsetArray
using Oceananigans.Architectures: GPU, CPU, AbstractArchitecture
using Oceananigans.Grids: AbstractGrid, RectilinearGrid, topology, Periodic, Bounded, Flat, Center, Face
using Oceananigans.Fields: Field, interior, location, instantiated_location
using Oceananigans.Models.NonhydrostaticModels: NonhydrostaticModel
using Oceananigans.TimeSteppers: RungeKutta3TimeStepper, QuasiAdamsBashforth2TimeStepper, tick!, tick_stage!, next_time, stage_Δt
using Oceananigans.Advection: WENO, Centered
using Oceananigans.TurbulenceClosures: ScalarDiffusivity, Smagorinsky, AnisotropicMinimumDissipation, VerticallyImplicitTimeDiscretization
using Oceananigans.BoundaryConditions: FieldBoundaryConditions, BoundaryCondition, Flux, Value, Gradient, Open, Periodic as PBC
using Oceananigans.DistributedComputations: Distributed, Partition
import Oceananigans.Architectures as AC
import Oceananigans.Fields as FD
import Oceananigans.Grids as GD
import Oceananigans.Solvers as SO
import Oceananigans.Utils as UT
import Oceananigans.BoundaryConditions as BC
import Oceananigans.TurbulenceClosures as TC
import Oceananigans.TimeSteppers as TS
import Oceananigans.Models.NonhydrostaticModels as NH
import Oceananigans.DistributedComputations as DC

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

# ---- architecture: B200 = GPU{B200Device}, exactly as ROCGPU = GPU{ROCBackend} (src/Architectures.jl:44-46) --------------------
# Every `arch isa GPU` branch of the reference (Fields/set!.jl:87,121; OutputWriters/fetch_output.jl:22; Solvers/*.jl) then takes
# the device path without further overrides.
mutable struct B200Device
    id  :: Int32
    ctx :: Ptr{Cvoid}    # ob_ctx*: stream, cuFFT plans, workspaces, communicator
    rank :: Int32
    world :: Int32
end
const B200 = GPU{B200Device}
function B200(id::Integer = 0)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    @ob ob_init (Int32, Ref{Ptr{Cvoid}}) Int32(id) ref
    dev = B200Device(Int32(id), ref[], Int32(0), Int32(1))
    finalizer(d -> ccall((:ob_shutdown, lib), Int32, (Ptr{Cvoid},), d.ctx), dev)
    return GPU(dev)
end
ctx(arch::B200) = arch.device.ctx

# ---- array type: owns device memory obtained from the library (ob_malloc / ob_free) ----------------------------------
mutable struct B200Array{T, N} <: AbstractArray{T, N}
    ptr  :: Ptr{Cvoid}
    dims :: NTuple{N, Int}
    arch :: B200
    function B200Array{T}(arch::B200, dims::NTuple{N, Int}) where {T, N}
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        @ob ob_malloc (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}) ctx(arch) prod(dims) * sizeof(T) ref   # zero-initialised
        a = new{T, N}(ref[], dims, arch)
        finalizer(x -> ccall((:ob_free, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(x.arch), x.ptr), a)  # thread-safe, no NCCL
        return a
    end
end
ftype(::Type{Float32}) = Int32(0)
ftype(::Type{Float64}) = Int32(1)
Base.size(a::B200Array) = a.dims
Base.pointer(a::B200Array) = a.ptr
Base.getindex(::B200Array, I...) = error("scalar indexing of a B200Array is disallowed (cf. allowscalar(false))")
Base.setindex!(::B200Array, v, I...) = error("scalar indexing of a B200Array is disallowed (cf. allowscalar(false))")
Base.similar(a::B200Array{T}, ::Type{S} = T, dims::Dims = size(a)) where {T, S} = B200Array{S}(a.arch, dims)
function Base.copyto!(dst::B200Array{T}, src::Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    @ob ob_memcpy_h2d (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) ctx(dst.arch) dst.ptr src sizeof(src)
    @ob ob_sync (Ptr{Cvoid},) ctx(dst.arch)      # `src` may be collected after return
    return dst
end
function Base.copyto!(dst::Array{T}, src::B200Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    @ob ob_memcpy_d2h (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) ctx(src.arch) dst src.ptr sizeof(dst)
    return dst
end
function Base.copyto!(dst::B200Array{T}, src::B200Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    @ob ob_memcpy_d2d (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t) ctx(dst.arch) dst.ptr src.ptr length(dst) * sizeof(T)
    return dst
end
# `parent(u) .= parent(v)` (Fields/set!.jl:228,240): whole-array identity broadcasts are copies; anything else is outside the path
Base.copyto!(dst::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{B200Array}}) = copyto!(dst, bc.args[1])
Base.copyto!(dst::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Array}}) = copyto!(dst, bc.args[1])
Base.copyto!(::B200Array, ::Base.Broadcast.Broadcasted) = error("B200Array supports fill!, copyto! and whole-array copies only; compute on a CPU field and set! it")
Base.Array(a::B200Array{T, N}) where {T, N} = copyto!(Array{T, N}(undef, size(a)), a)
function Base.fill!(a::B200Array{T}, v) where T
    @ob ob_fill (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Float64) ctx(a.arch) a.ptr length(a) ftype(T) Float64(v)
    return a
end
function Base.any(::typeof(isnan), a::B200Array{T}) where T       # Diagnostics/nan_checker.jl
    flag = Ref{Int32}(0)
    @ob ob_any_nan (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Ref{Int32}) ctx(a.arch) a.ptr length(a) ftype(T) flag
    return flag[] != 0
end

# ---- src/Architectures.jl:60-132 and Grids/zeros_and_ones.jl:8 for the new architecture ----------------------------------
AC.device(a::B200) = a.device
AC.device!(a::B200) = nothing                      # the library selects the context's device on every call
AC.device!(a::B200, i) = nothing
AC.ndevices(::B200) = (n = Ref{Int32}(0); @ob(ob_device_count, (Ref{Int32},), n); Int(n[]))
AC.synchronize(a::B200) = @ob ob_sync (Ptr{Cvoid},) ctx(a)
UT.sync_device!(a::B200) = AC.synchronize(a)
AC.array_type(::B200) = B200Array
AC.architecture(a::B200Array) = a.arch
AC.architecture(::Type{<:B200Array}) = error("the architecture of a B200Array type needs an instance (it carries the context)")
AC.on_architecture(arch::B200, a::Array{T}) where T = copyto!(B200Array{T}(arch, size(a)), a)
AC.on_architecture(arch::B200, a::BitArray) = AC.on_architecture(arch, Array{Bool}(a))
AC.on_architecture(arch::B200, a::StepRangeLen) = a
AC.on_architecture(::CPU, a::B200Array) = Array(a)
AC.on_architecture(::B200, a::B200Array) = a
AC.unified_array(::B200, a) = a
AC.device_copy_to!(dst::B200Array, src::B200Array; kw...) = copyto!(dst, src)
AC.unsafe_free!(a::B200Array) = (ccall((:ob_free, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(a.arch), a.ptr); a.ptr = C_NULL; nothing)
AC.convert_to_device(::B200, args) = args            # no kernel arguments are ever converted: no KA kernel is launched
Base.zeros(arch::B200, FT, N...) = B200Array{FT}(arch, Tuple(Int.(N)))       # Grids/zeros_and_ones.jl:8
BC.validate_boundary_condition_architecture(::B200Array, ::B200, bc, side) = nothing
BC.validate_boundary_condition_architecture(::Array, ::B200, bc, side) =
    throw(ArgumentError("$side $bc must use `B200Array` rather than `Array` on the B200 architecture (on_architecture(arch, array))"))
Oceananigans.Advection.default_weno_weight_computation(::B200) = UT.BackendOptimizedDivision
# nothing may silently fall back to KernelAbstractions on this architecture:
UT.launch!(::B200, args...; kw...) = error("no KernelAbstractions path on B200: this operation is outside the accelerated NonhydrostaticModel hot path")

# ---- POD descriptors: field-for-field mirrors of include/ocean_b200.h (checked by tests/test_abi.py) -------------------------
struct ObGridDesc
    float_type::Int32; N::NTuple{3, Int32}; H::NTuple{3, Int32}; topology::NTuple{3, Int32}
    L::NTuple{3, Float64}; d::NTuple{3, Float64}
    dzf_host::Ptr{Cvoid}; dzc_host::Ptr{Cvoid}; n_dzf::Int32; n_dzc::Int32
end
struct ObBcDesc; kind::NTuple{6, Int32}; value::NTuple{6, Float64}; end
struct ObClosureDesc
    kind::Int32; nu::Float64; kappa::NTuple{8, Float64}; cs::Float64; lilly::Int32; cb::Float64; Pr::NTuple{8, Float64}
    Cnu::Float64; Ckappa::NTuple{8, Float64}; amd_has_cb::Int32; vertically_implicit::Int32
    dynamic::Int32; averaging_dims::Int32; minimum_numerator::Float64
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
const B200Grid = AbstractGrid{<:Any, <:Any, <:Any, <:Any, <:B200}
const DistB200 = Distributed{<:B200}
const AnyB200Grid = AbstractGrid{<:Any, <:Any, <:Any, <:Any, <:Union{B200, DistB200}}
const B200Field = Field{<:Any, <:Any, <:Any, <:Any, <:AnyB200Grid}
b200(arch::B200) = arch
b200(arch::DistB200) = arch.child_architecture

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
    return desc, (dzf, dzc)   # keep the host arrays alive across the call that receives `desc`
end
# under Distributed the model descriptor carries the GLOBAL grid (the library keeps this rank's equal x slab)
grid_desc(grid::RectilinearGrid{<:Any, <:Any, <:Any, <:Any, <:Any, <:Any, <:Any, <:Any, <:Any, <:DistB200}) =
    grid_desc(DC.reconstruct_global_grid(grid))

# ---- translation of the model description; anything the ABI cannot express throws (no CPU fallback) -----------------
unsupported(what) = throw(ArgumentError("B200: $what is outside the accelerated NonhydrostaticModel path (SURVEY.md §2)"))

bc_kind(::Nothing) = (Int32(0), 0.0)
function bc_kind(bc::BoundaryCondition)
    c = bc.classification
    c isa PBC && return (Int32(1), 0.0)
    c isa DC.DistributedCommunication && return (Int32(6), 0.0)
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
sides(bcs::FieldBoundaryConditions) = (bcs.west, bcs.east, bcs.south, bcs.north, bcs.bottom, bcs.top)
function bc_desc(bcs::FieldBoundaryConditions)
    ks = map(bc_kind, sides(bcs))
    return ObBcDesc(ntuple(i -> ks[i][1], 6), ntuple(i -> ks[i][2], 6))
end
bc_desc(::Nothing) = ObBcDesc(ntuple(_ -> Int32(0), 6), ntuple(_ -> 0.0, 6))
bc_array_ptrs(bcs::FieldBoundaryConditions) =
    [(bc isa BoundaryCondition && bc.condition isa B200Array) ? pointer(bc.condition) : C_NULL for bc in sides(bcs)]

pad8(t) = ntuple(i -> i <= length(t) ? Float64(t[i]) : 0.0, 8)
const NO_CLOSURE = ObClosureDesc(0, 0.0, pad8(()), 0.0, 0, 0.0, pad8(()), 0.0, pad8(()), 0, 0, 0, 0, 0.0)
closure_desc(c::ScalarDiffusivity, names) =
    (c.ν isa Number && all(κ -> κ isa Number, values(c.κ))) ?
        ObClosureDesc(1, Float64(c.ν), pad8(values(c.κ)), 0.0, 0, 0.0, pad8(()), 0.0, pad8(()), 0,
                      Int32(TC.time_discretization(c) isa VerticallyImplicitTimeDiscretization), 0, 0, 0.0) : unsupported("a function-valued diffusivity")
function closure_desc(c::Smagorinsky, names)
    coeff = c.coefficient
    vi = Int32(TC.time_discretization(c) isa VerticallyImplicitTimeDiscretization)
    if coeff isa TC.Smagorinskys.DynamicCoefficient   # DynamicSmagorinsky: directional averaging only, updated every iteration
        avg = coeff.averaging
        (avg isa TC.Smagorinskys.LagrangianAveraging) && unsupported("DynamicSmagorinsky with LagrangianAveraging")
        (coeff.schedule isa Oceananigans.Utils.IterationInterval && coeff.schedule.interval == 1 && coeff.schedule.offset == 0) ||
            unsupported("DynamicCoefficient schedules other than IterationInterval(1)")
        dims = avg isa Colon ? (1, 2, 3) : avg
        mask = Int32(sum(1 << (d - 1) for d in dims))
        return ObClosureDesc(2, 0.0, pad8(()), 0.0, 0, 0.0, pad8(values(c.Pr)), 0.0, pad8(()), 0, vi, 1, mask, Float64(coeff.minimum_numerator))
    end
    lilly = !(coeff isa Number)
    cs = lilly ? coeff.smagorinsky : coeff
    cb = lilly ? coeff.reduction_factor : 0.0
    return ObClosureDesc(2, 0.0, pad8(()), Float64(cs), Int32(lilly), Float64(cb), pad8(values(c.Pr)), 0.0, pad8(()), 0, vi, 0, 0, 0.0)
end
closure_desc(c::AnisotropicMinimumDissipation, names) =
    ObClosureDesc(3, 0.0, pad8(()), 0.0, 0, c.Cb === nothing ? 0.0 : Float64(c.Cb), pad8(()), Float64(c.Cν), pad8(values(c.Cκ)), Int32(c.Cb !== nothing),
                  Int32(TC.time_discretization(c) isa VerticallyImplicitTimeDiscretization), 0, 0, 0.0)
closure_desc(c, names) = unsupported("closure $(typeof(c))")
closure_tuple(model) = model.closure === nothing ? () : model.closure isa Tuple ? model.closure : (model.closure,)

function model_desc(model::NonhydrostaticModel)
    grid = model.grid
    gd, keep = grid_desc(grid)
    adv = model.advection.momentum   # the shim requires one scheme for momentum and tracers
    all(a -> a === adv || a == adv, values(model.advection)) || unsupported("different advection schemes for momentum and tracers")
    kind, order = adv isa WENO ? (Int32(2), 2 * Oceananigans.Advection.required_halo_size_x(adv) - 1) :
                  adv isa Centered ? (Int32(1), 2 * Oceananigans.Advection.required_halo_size_x(adv)) : unsupported("advection $(typeof(adv))")
    closures = closure_tuple(model)
    length(closures) <= 4 || unsupported("more than 4 closures")
    names = keys(model.tracers)
    length(names) <= 8 || unsupported("more than 8 tracers")
    cds = ntuple(i -> i <= length(closures) ? closure_desc(closures[i], names) : NO_CLOSURE, 4)
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
    all(f -> f === Oceananigans.Forcings.zeroforcing, values(model.forcing)) || unsupported("user forcing functions")
    (model.stokes_drift === nothing && model.particles === nothing && model.biogeochemistry === nothing && model.free_surface === nothing) ||
        unsupported("Stokes drift / particles / biogeochemistry / free surface")
    ts = model.timestepper
    tr_bcs = ntuple(i -> i <= length(names) ? bc_desc(model.tracers[i].boundary_conditions) : bc_desc(nothing), 8)
    none4 = ntuple(i -> bc_desc(nothing), 4); none48 = ntuple(i -> ntuple(j -> bc_desc(nothing), 8), 4)   # closure-field defaults are filled in by the library
    pHY = model.pressures.pHY′
    weno_div = (adv isa WENO && !(adv.weight_computation isa Type{<:UT.NormalDivision})) ? Int32(1) : Int32(0)
    desc = ObModelDesc(gd, kind, Int32(order), weno_div, Int32(length(closures)), cds, bk, ib, iT, iS, g, α, β,
                       Int32(cor !== nothing), cor === nothing ? 0.0 : Float64(cor.f), Int32(length(names)),
                       ts isa RungeKutta3TimeStepper ? Int32(0) : Int32(1), ts isa RungeKutta3TimeStepper ? 0.1 : Float64(ts.χ),
                       Int32(pHY !== nothing), bc_desc(model.velocities.u.boundary_conditions), bc_desc(model.velocities.v.boundary_conditions),
                       bc_desc(model.velocities.w.boundary_conditions), bc_desc(model.pressures.pNHS.boundary_conditions),
                       pHY === nothing ? bc_desc(nothing) : bc_desc(pHY.boundary_conditions), tr_bcs, none4, none48)
    return desc, keep
end

# ---- model handle: created lazily at the first hot-path call, cached per model ----------------------------------------------
const B200Model{TS} = NonhydrostaticModel{TS, <:Any, <:Union{B200, DistB200}}
const HANDLES = IdDict{Any, Ptr{Cvoid}}()
handle(model) = get!(() -> create_handle(model), HANDLES, model)
mctx(model) = ctx(b200(model.architecture))

function bind!(h, id::Integer, f)   # ob_model_bind_field(model, field_id, device pointer of parent(f))
    @ob ob_model_bind_field (Ptr{Cvoid}, Int32, Ptr{Cvoid}) h Int32(id) pointer(parent(f))
end

function create_handle(model)
    desc, keep = model_desc(model)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep @ob ob_model_create (Ptr{Cvoid}, Ref{ObModelDesc}, Ref{Ptr{Cvoid}}) mctx(model) Ref(desc) ref
    h = ref[]
    bind!(h, 0, model.velocities.u); bind!(h, 1, model.velocities.v); bind!(h, 2, model.velocities.w)
    bind!(h, 3, model.pressures.pNHS)
    model.pressures.pHY′ === nothing || bind!(h, 4, model.pressures.pHY′)
    for (t, c) in enumerate(model.tracers); bind!(h, 16 + t - 1, c); end
    Gⁿ, G⁻ = model.timestepper.Gⁿ, model.timestepper.G⁻
    for (n, (a, b)) in enumerate(zip(Gⁿ, G⁻)); bind!(h, 32 + n - 1, a); bind!(h, 48 + n - 1, b); end
    for (m, c) in enumerate(closure_tuple(model))
        K = model.closure isa Tuple ? model.closure_fields[m] : model.closure_fields
        c isa ScalarDiffusivity && continue
        bind!(h, 64 + m - 1, K.νₑ)
        if c isa AnisotropicMinimumDissipation
            for (t, κ) in enumerate(K.κₑ); bind!(h, 80 + (m - 1) * 8 + t - 1, κ); end
        end
    end
    # array-valued conditions: condition[i, j] over the boundary plane, already a device array (on_architecture(::B200, ...))
    for (id, f) in ((0, model.velocities.u), (1, model.velocities.v), (2, model.velocities.w), ((16 + t - 1, c) for (t, c) in enumerate(model.tracers))...)
        for (side, p) in enumerate(bc_array_ptrs(f.boundary_conditions))
            p == C_NULL && continue
            @ob ob_model_set_bc_array (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}) h Int32(id) Int32(side - 1) p
        end
    end
    finalizer(_ -> (ccall((:ob_model_destroy, lib), Int32, (Ptr{Cvoid},), h); nothing), model.timestepper)
    return h
end

# ---- model construction hooks (nonhydrostatic_model.jl:124-313) ----------------------------------------------------------------
# pressure_solver: the reference would call FFTBasedPoissonSolver(grid) -> plan_forward_transform on a device array.  The
# library owns the transforms; the model carries a handle-wrapping solver so that `model.pressure_solver` exists and `solve!` works.
mutable struct B200PoissonSolver{G}
    grid   :: G
    handle :: Ptr{Cvoid}
end
function B200PoissonSolver(grid)
    desc, keep = grid_desc(grid)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep @ob ob_solver_create (Ptr{Cvoid}, Ref{ObGridDesc}, Ref{Ptr{Cvoid}}) ctx(b200(GD.architecture(grid))) Ref(desc) ref
    s = B200PoissonSolver(grid, ref[])
    finalizer(x -> (ccall((:ob_solver_destroy, lib), Int32, (Ptr{Cvoid},), x.handle); nothing), s)
    return s
end
NH.nonhydrostatic_pressure_solver(::B200, grid::NH.XYZRegularRG, ::Nothing) = B200PoissonSolver(grid)
NH.nonhydrostatic_pressure_solver(::B200, grid::NH.GridWithFourierTridiagonalSolver, ::Nothing) = B200PoissonSolver(grid)
NH.nonhydrostatic_pressure_solver(::DistB200, grid::NH.XYZRegularRG, ::Nothing) = B200PoissonSolver(grid)
NH.nonhydrostatic_pressure_solver(::DistB200, grid::NH.GridWithFourierTridiagonalSolver, ::Nothing) = B200PoissonSolver(grid)
NH.nonhydrostatic_pressure_solver(::Union{B200, DistB200}, grid, ::Nothing) = unsupported("a grid without an FFT-based pressure solver")
NH.nonhydrostatic_pressure_solver(::Union{B200, DistB200}, grid, free_surface) = unsupported("a free surface")
"solve!(ϕ, solver, rhs): rhs and ϕ are (Nx, Ny, Nz) device arrays without halos (fft_based_poisson_solver.jl:94-124)"
function SO.solve!(ϕ::B200Array, solver::B200PoissonSolver, rhs::B200Array)
    @ob ob_poisson_solve (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}) solver.handle pointer(rhs) pointer(ϕ)
    return ϕ
end
# the vertically-implicit solver object the TimeStepper asks for (implicit_diffusion_solver): the library solves the columns inside
# ob_rk3_substep / ob_ab2_step, so the model only needs a placeholder
struct B200ImplicitSolver end
TC.implicit_diffusion_solver(::VerticallyImplicitTimeDiscretization, ::AnyB200Grid) = B200ImplicitSolver()

# ---- fields: set!, halo fills -------------------------------------------------------------------------------------------------------
"host staging of an interior assignment: parent -> host, write the interior window, host -> parent"
function set_interior_from_host!(u::B200Field, vals::AbstractArray)
    h = Array(parent(u))
    hv = OffsetArray(h, u.data.offsets...)
    interior(hv, instantiated_location(u), u.grid, u.indices) .= reshape(vals, size(u)...)
    copyto!(parent(u), h)
    return u
end
FD.set_to_array!(u::B200Field, a::Array) = set_interior_from_host!(u, a)
FD.set_to_array!(u::B200Field, a::B200Array) = set_interior_from_host!(u, Array(a))
function FD.copy_to_field!(u::B200Field, v::Field)     # set!(u, cpu_field) at the end of set_to_function! (Fields/set!.jl:121)
    if size(parent(u)) == size(parent(v))
        src = parent(v)
        copyto!(parent(u), src isa B200Array ? src : Array(src))     # halos travel along, as the reference attempts
    else
        set_interior_from_host!(u, Array(interior(v)))
    end
    return u
end

loc_id(::Face) = Int32(1); loc_id(::Center) = Int32(0); loc_id(::Nothing) = Int32(0)
"fill_halo_regions!(c, bcs, indices, loc, grid, args...) for any B200Array-backed field data (fill_halo_regions.jl:20-38)"
function BC.fill_halo_regions!(c::OffsetArray{<:Any, <:Any, <:B200Array}, bcs::FieldBoundaryConditions, indices, loc, grid::B200Grid, args...;
                               fill_normal_flow_bcs = true, kwargs...)
    all(i -> i isa Colon, indices) || unsupported("halo fills of windowed fields")
    desc, keep = grid_desc(grid)
    ptrs = bc_array_ptrs(bcs)
    locs = Int32[loc_id(loc[1]), loc_id(loc[2]), loc_id(loc[3])]
    GC.@preserve keep ptrs locs @ob ob_fill_halo_array (Ptr{Cvoid}, Ref{ObGridDesc}, Ptr{Cvoid}, Ptr{Int32}, Ref{ObBcDesc}, Ptr{Ptr{Cvoid}}, Int32) ctx(GD.architecture(grid)) Ref(desc) pointer(parent(c)) locs Ref(bc_desc(bcs)) ptrs Int32(fill_normal_flow_bcs)
    return nothing
end
# distributed fields exchange their x halos inside the library; the shim reaches them through the model (field ids)
function fill_model_halo!(model::B200Model, field_id::Integer; fill_normal_flow_bcs = true)
    @ob ob_fill_halo (Ptr{Cvoid}, Int32, Int32) handle(model) Int32(field_id) Int32(fill_normal_flow_bcs)
end

# ---- hot-path method overrides (SURVEY.md §8b item 3) ----------------------------------------------------------------
TS.update_state!(model::B200Model, callbacks = []; kw...) =
    (@ob ob_update_state (Ptr{Cvoid},) handle(model); foreach(c -> c.callsite isa TS.UpdateStateCallsite && c(model), callbacks); nothing)
TS.compute_tendencies!(model::B200Model, callbacks = []) =
    (@ob ob_compute_tendencies (Ptr{Cvoid},) handle(model); foreach(c -> c.callsite isa TS.TendencyCallsite && c(model), callbacks); nothing)
TS.compute_flux_bc_tendencies!(model::B200Model) = nothing      # folded into ob_rk3_substep / ob_ab2_step (flux_bc_kernel runs first there)
TS.cache_previous_tendencies!(model::B200Model) = @ob ob_cache_tendencies (Ptr{Cvoid},) handle(model)
NH.compute_auxiliaries!(model::B200Model; kw...) = (TC.compute_closure_fields!(model.closure_fields, model.closure, model); NH.update_hydrostatic_pressure!(model); nothing)
TC.compute_closure_fields!(closure_fields, closure, model::B200Model; kw...) = @ob ob_compute_closure_fields (Ptr{Cvoid},) handle(model)
NH.update_hydrostatic_pressure!(model::B200Model; kw...) = @ob ob_update_hydrostatic_pressure (Ptr{Cvoid},) handle(model)
NH.compute_pressure_correction!(model::B200Model, Δt) = @ob ob_compute_pressure_correction (Ptr{Cvoid}, Float64) handle(model) Float64(Δt)
NH.make_pressure_correction!(model::B200Model, Δt) = @ob ob_make_pressure_correction (Ptr{Cvoid}, Float64) handle(model) Float64(Δt)
TC.initialize_closure_fields!(closure_fields, closure, model::B200Model) = nothing   # nothing to initialise for the closures in scope

"rk3_substep!(model, Δt, γ, ζ, callbacks): update + implicit columns + pressure correction of one stage (nonhydrostatic_rk3_substep.jl:31-63)"
function TS.rk3_substep!(model::B200Model, Δt, γ, ζ, callbacks)
    @ob ob_rk3_substep (Ptr{Cvoid}, Float64, Float64, Float64, Int32) handle(model) Float64(Δt) Float64(γ) Float64(something(ζ, 0)) Int32(!isnothing(ζ))
    return nothing
end
"ab2_step!(model, Δt, callbacks) (nonhydrostatic_ab2_step.jl:10-57); χ = -1/2 on Euler steps (quasi_adams_bashforth_2.jl:104-110)"
function TS.ab2_step!(model::B200Model, Δt, callbacks)
    @ob ob_ab2_step (Ptr{Cvoid}, Float64, Float64) handle(model) Float64(Δt) Float64(model.timestepper.χ)
    return nothing
end

# With the methods above the reference's own time_step! drivers (runge_kutta_3.jl:103-168, quasi_adams_bashforth_2.jl:90-126) run
# unchanged on a B200 model: every call they make lands on an override.  The two methods below are the FAST path -- one C call per
# step -- taken when no callbacks are registered; clock handling is the reference's (tick_stage!, a-priori tⁿ⁺¹).
function TS.time_step!(model::B200Model{<:RungeKutta3TimeStepper}, Δt; callbacks = [])
    isempty(callbacks) || return invoke(TS.time_step!, Tuple{Oceananigans.AbstractModel{<:RungeKutta3TimeStepper}, Any}, model, Δt; callbacks)
    first = model.clock.iteration == 0
    ts = model.timestepper
    tⁿ⁺¹ = next_time(model.clock, Δt)
    @ob ob_time_step_rk3 (Ptr{Cvoid}, Float64, Int32) handle(model) Float64(Δt) Int32(first)
    tick_stage!(model.clock, stage_Δt(Δt, ts.γ¹, nothing))
    tick_stage!(model.clock, stage_Δt(Δt, ts.γ², ts.ζ²))
    tick_stage!(model.clock, Oceananigans.Units.time_difference_seconds(tⁿ⁺¹, model.clock.time), Δt)
    return nothing
end
function TS.time_step!(model::B200Model{<:QuasiAdamsBashforth2TimeStepper}, Δt; callbacks = [], euler = false)
    isempty(callbacks) || return invoke(TS.time_step!, Tuple{Oceananigans.AbstractModel{<:QuasiAdamsBashforth2TimeStepper}, Any}, model, Δt; callbacks, euler)
    euler = euler || (Δt != model.clock.last_Δt)
    @ob ob_time_step_ab2 (Ptr{Cvoid}, Float64, Int32, Int32) handle(model) Float64(Δt) Int32(euler) Int32(model.clock.iteration == 0)
    tick!(model.clock, Δt)
    return nothing
end

# TimeStepWizard (cell_advection_timescale.jl:14-35): device min-reduction
function Oceananigans.Advection.cell_advection_timescale(model::B200Model)
    τ = Ref{Float64}(0)
    @ob ob_cell_advection_timescale (Ptr{Cvoid}, Ref{Float64}) handle(model) τ
    return τ[]
end

# ---- Distributed(B200(); partition = Partition(R)): one process per GPU, slab-x ---------------------------------------------------
# The reference constructor (distributed_architectures.jl:240-305) initialises MPI and builds the rank connectivity; that host
# logic is reused as is.  The device side -- NCCL communicator, CUDA-IPC halo staging, transposes -- lives in the library: rank 0
# creates the id, MPI broadcasts it (exactly nccl_communicator.jl:25-63), every rank calls ob_dist_init on its context.
function DC.Distributed(child::B200; partition = nothing, kwargs...)
    arch = invoke(DC.Distributed, Tuple{AbstractArchitecture}, child; partition, kwargs...)
    (arch.ranks[2] == 1 && arch.ranks[3] == 1) || unsupported("pencil partitions (slab-x only)")
    MPI = DC.MPI
    id = zeros(UInt8, 128)
    arch.local_rank == 0 && @ob ob_dist_unique_id (Ptr{UInt8},) id
    MPI.Bcast!(id, 0, arch.communicator)
    @ob ob_dist_init (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}) ctx(child) Int32(arch.local_rank) Int32(prod(arch.ranks)) id
    child.device.rank, child.device.world = Int32(arch.local_rank), Int32(prod(arch.ranks))
    return arch
end
# halo exchange and transposes of a distributed B200 field never go through MPI/KA: they belong to the model's library handle
DC.distributed_fill_halo_event!(c, kernel!, bcs, loc, grid::AbstractGrid{<:Any, <:Any, <:Any, <:Any, <:DistB200}, buffers, args...; kwargs...) =
    error("distributed halos on B200 are exchanged by the library: call fill_halo_regions! through the model (update_state!, fill_model_halo!)")
DC.transpose_z_to_y!(::DC.TransposableField{<:B200Field}) = error("distributed transposes on B200 happen inside ob_poisson_solve")
DC.transpose_y_to_x!(::DC.TransposableField{<:B200Field}) = error("distributed transposes on B200 happen inside ob_poisson_solve")
DC.transpose_x_to_y!(::DC.TransposableField{<:B200Field}) = error("distributed transposes on B200 happen inside ob_poisson_solve")
DC.transpose_y_to_z!(::DC.TransposableField{<:B200Field}) = error("distributed transposes on B200 happen inside ob_poisson_solve")

end # module
