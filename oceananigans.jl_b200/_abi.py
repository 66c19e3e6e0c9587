"""ctypes binding of the C ABI declared in include/ocean_b200.h.

This is the Python twin of the `ccall` stubs shown in INTEGRATION.md (the Julia extension binds exactly the
same symbols with the same POD structs).  There is no CPU fallback: if libocean_b200.so is missing or no
sm_100 device is present every entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCEAN_B200_LIB") or os.path.join(HERE, "libocean_b200.so")   # same override as the Julia shim

OB_MAX_TRACERS = 8
OB_MAX_CLOSURES = 4
OB_F32, OB_F64 = 0, 1
OB_PERIODIC, OB_BOUNDED, OB_FLAT = 0, 1, 2
(OB_BC_NONE, OB_BC_PERIODIC, OB_BC_FLUX, OB_BC_VALUE, OB_BC_GRADIENT, OB_BC_IMPENETRABLE, OB_BC_COMMUNICATION) = range(7)
OB_ADV_NONE, OB_ADV_CENTERED, OB_ADV_WENO = 0, 1, 2
OB_CLOSURE_SCALAR_DIFFUSIVITY, OB_CLOSURE_SMAGORINSKY, OB_CLOSURE_AMD = 1, 2, 3
OB_BUOYANCY_NONE, OB_BUOYANCY_TRACER, OB_BUOYANCY_LINEAR_SEAWATER = 0, 1, 2
OB_RK3, OB_AB2 = 0, 1
OB_DIV_EXACT, OB_DIV_RCP_NEWTON = 0, 1
OB_OPT_TENDENCY_KERNEL = 1
OB_OPT_FUSE_PROJECTION = 2
OB_OPT_OVERLAP_HALO = 3
OB_OPT_VECTOR_STREAMS = 4
OB_FIELD_U, OB_FIELD_V, OB_FIELD_W, OB_FIELD_PNHS, OB_FIELD_PHY = 0, 1, 2, 3, 4
OB_FIELD_TRACER0, OB_FIELD_GN0, OB_FIELD_GM0, OB_FIELD_NUE0, OB_FIELD_KAPPAE0 = 16, 32, 48, 64, 80


class GridDesc(C.Structure):
    _fields_ = [
        ("float_type", C.c_int32), ("N", C.c_int32 * 3), ("H", C.c_int32 * 3), ("topology", C.c_int32 * 3),
        ("L", C.c_double * 3), ("d", C.c_double * 3),
        ("dzf_host", C.c_void_p), ("dzc_host", C.c_void_p), ("n_dzf", C.c_int32), ("n_dzc", C.c_int32),
    ]


class BcDesc(C.Structure):
    _fields_ = [("kind", C.c_int32 * 6), ("value", C.c_double * 6)]


class ClosureDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nu", C.c_double), ("kappa", C.c_double * OB_MAX_TRACERS),
        ("cs", C.c_double), ("lilly", C.c_int32), ("cb", C.c_double), ("Pr", C.c_double * OB_MAX_TRACERS),
        ("Cnu", C.c_double), ("Ckappa", C.c_double * OB_MAX_TRACERS), ("amd_has_cb", C.c_int32), ("vertically_implicit", C.c_int32),
        ("dynamic", C.c_int32), ("averaging_dims", C.c_int32), ("minimum_numerator", C.c_double),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("grid", GridDesc),
        ("advection_kind", C.c_int32), ("advection_order", C.c_int32), ("weno_division", C.c_int32),
        ("n_closures", C.c_int32), ("closures", ClosureDesc * OB_MAX_CLOSURES),
        ("buoyancy_kind", C.c_int32), ("buoyancy_tracer", C.c_int32),
        ("temperature_tracer", C.c_int32), ("salinity_tracer", C.c_int32),
        ("g", C.c_double), ("thermal_expansion", C.c_double), ("haline_contraction", C.c_double),
        ("has_coriolis", C.c_int32), ("f", C.c_double),
        ("n_tracers", C.c_int32), ("stepper", C.c_int32), ("chi", C.c_double),
        ("has_hydrostatic_pressure", C.c_int32),
        ("bcs_u", BcDesc), ("bcs_v", BcDesc), ("bcs_w", BcDesc), ("bcs_p", BcDesc), ("bcs_phy", BcDesc),
        ("bcs_tracer", BcDesc * OB_MAX_TRACERS), ("bcs_nue", BcDesc * OB_MAX_CLOSURES),
        ("bcs_kappae", (BcDesc * OB_MAX_TRACERS) * OB_MAX_CLOSURES),
    ]


class OceanB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libocean_b200 error %d: %s" % (code, msg))
        self.code = code


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_i32, _i64, _dbl, _sz = C.c_int32, C.c_int64, C.c_double, C.c_size_t

# name -> argtypes (all return int32 status unless listed in _STR)
PROTOTYPES = {
    "ob_init": [_i32, _PP], "ob_shutdown": [_P], "ob_device_count": [C.POINTER(_i32)], "ob_sync": [_P],
    "ob_fp64_peak": [_P, C.POINTER(_dbl)],
    "ob_timer_start": [_P], "ob_timer_stop": [_P, C.POINTER(_dbl)],
    "ob_malloc": [_P, _sz, _PP], "ob_free": [_P, _P], "ob_malloc_host": [_P, _sz, _PP], "ob_free_host": [_P, _P],
    "ob_memcpy_h2d": [_P, _P, _P, _sz], "ob_memcpy_d2h": [_P, _P, _P, _sz], "ob_memcpy_d2d": [_P, _P, _P, _sz],
    "ob_memcpy_d2h_async": [_P, _P, _P, _sz], "ob_stream_wait": [_P, _P],
    "ob_fill": [_P, _P, _sz, _i32, _dbl], "ob_any_nan": [_P, _P, _sz, _i32, C.POINTER(_i32)],
    "ob_cell_advection_timescale": [_P, C.POINTER(_dbl)],
    "ob_model_create": [_P, C.POINTER(ModelDesc), _PP], "ob_model_destroy": [_P],
    "ob_model_bind_field": [_P, _i32, _P], "ob_model_set_bc_array": [_P, _i32, _i32, _P], "ob_fill_halo": [_P, _i32, _i32],
    "ob_fill_halo_array": [_P, C.POINTER(GridDesc), _P, C.POINTER(_i32 * 3), C.POINTER(BcDesc), _P, _i32],
    "ob_update_state": [_P], "ob_compute_tendencies": [_P], "ob_compute_closure_fields": [_P],
    "ob_update_hydrostatic_pressure": [_P],
    "ob_rk3_substep": [_P, _dbl, _dbl, _dbl, _i32], "ob_ab2_step": [_P, _dbl, _dbl], "ob_cache_tendencies": [_P],
    "ob_compute_pressure_correction": [_P, _dbl], "ob_make_pressure_correction": [_P, _dbl],
    "ob_time_step_rk3": [_P, _dbl, _i32], "ob_time_step_ab2": [_P, _dbl, _i32, _i32],
    "ob_model_set_option": [_P, _i32, _i32],
    "ob_launch_count": [_P, C.POINTER(_i64)], "ob_enable_timing": [_P, _i32], "ob_phase_count": [C.POINTER(_i32)],
    "ob_phase_time_ms": [_P, _i32, C.POINTER(_dbl), C.POINTER(_i64)], "ob_reset_timing": [_P],
    "ob_solver_create": [_P, C.POINTER(GridDesc), _PP], "ob_solver_destroy": [_P],
    "ob_poisson_solve": [_P, _P, _P],
    "ob_batched_tridiagonal_solve": [_P, _i32, _i32, _i32, _i32, _i32, _P, _P, _P, _P, _P, _P],
    "ob_dist_unique_id": [_P], "ob_dist_init": [_P, _i32, _i32, _P], "ob_dist_finalize": [_P],
}
_STR = {"ob_last_error": [], "ob_phase_name": [_i32]}

_lib = None


def lib():
    """Load libocean_b200.so (built in-tree by oceananigans.jl_b200/build.py).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OceanB200Error(-4, "%s not found: run `python -m __graft_entry__` / build.py (no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int32
        for name, args in _STR.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_char_p
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise OceanB200Error(status, lib().ob_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(lib(), name)(*args))
