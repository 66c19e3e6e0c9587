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
using Oceananigans.TurbulenceClosures: ScalarDiffusivity, Smagorinsky, AnisotropicMinimumDissipation
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
    Cnu::Float64; Ckappa::NTuple{8, Float64}; amd_has_cb::Int32
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

# bc_desc, closure_desc, model_desc: translate FieldBoundaryConditions / closures / buoyancy / coriolis into the POD
# structs, throwing ArgumentError for anything the ABI cannot express (function-valued BCs and forcings, background
# fields, immersed boundaries, other closures / equations of state) -- see models.py for the executable twin of this
# validation logic.

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

# fill_halo_regions!(field) for fields that belong to a B200 model: ob_fill_halo(model, field_id, fill_normal_flow_bcs)
# TimeStepWizard: cell_advection_timescale(model::B200Model) -> ob_cell_advection_timescale(handle, Ref{Float64})

end # module
