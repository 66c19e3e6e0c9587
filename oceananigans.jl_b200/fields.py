"""Fields and boundary conditions (host-side mirror).

Reference: src/Fields/field.jl:21-36 (Field = OffsetArray over a padded parent), src/Grids/new_data.jl:11-74
(parent size (N+2H) per dim, +1 for Face on Bounded, no halo on Flat), src/Fields/set!.jl:63-160,
src/BoundaryConditions/field_boundary_conditions.jl:15-36 (defaults).  Field memory is device memory obtained
from the library (ob_malloc) -- the Python object only holds the pointer, exactly like `B200Array` in the
Julia extension.
"""
import ctypes as C

import numpy as np

from . import _abi


class BoundaryCondition:
    def __init__(self, kind, value=None):
        self.kind, self.value = kind, value

    def __repr__(self):
        return "%s(%r)" % (self.kind, self.value)


def FluxBoundaryCondition(value):
    """`value`: a number, or an array over the boundary plane -- numpy shape (n_slow, n_fast) of the field's interior extent in
    the two tangential dimensions, e.g. (ny, nx) for top/bottom (BoundaryCondition with an AbstractArray condition)"""
    return BoundaryCondition("Flux", value)


def ValueBoundaryCondition(value):
    return BoundaryCondition("Value", value)


def GradientBoundaryCondition(value):
    return BoundaryCondition("Gradient", value)


def OpenBoundaryCondition(value=None):
    if value is not None:
        raise _abi.OceanB200Error(-3, "OpenBoundaryCondition with a prescribed value is outside the B200 hot path")
    return BoundaryCondition("Impenetrable", None)


SIDES = ("west", "east", "south", "north", "bottom", "top")


class FieldBoundaryConditions:
    """FieldBoundaryConditions(; west, east, south, north, bottom, top): unspecified sides take the defaults
    of field_boundary_conditions.jl:15-36 when the model regularises them."""

    def __init__(self, **kw):
        for k in kw:
            if k not in SIDES:
                raise ValueError("unknown side %r" % k)
        self.sides = dict(kw)


def default_bc(topo, loc, auxiliary=False):
    if topo == _abi.OB_PERIODIC:
        return BoundaryCondition("Periodic")
    if topo == _abi.OB_FLAT:
        return None
    if loc == "c":
        return BoundaryCondition("Flux", None)  # NoFluxBoundaryCondition
    return None if auxiliary else BoundaryCondition("Impenetrable")


def regularize_bcs(grid, loc, user=None, auxiliary=False):
    out = {}
    user = user.sides if isinstance(user, FieldBoundaryConditions) else (user or {})
    for s, side in enumerate(SIDES):
        d = s // 2
        bc = default_bc(grid.topo[d], loc[d], auxiliary)
        if side in user and user[side] is not None:
            ub = user[side]
            if callable(ub.value):
                raise _abi.OceanB200Error(-3, "function-valued boundary conditions cannot cross the C ABI (SURVEY.md §2)")
            if grid.topo[d] != _abi.OB_BOUNDED:
                raise ValueError("cannot set a %s boundary condition in a non-Bounded direction" % side)
            bc = ub
        out[side] = bc
    return out


_KIND = {"Periodic": _abi.OB_BC_PERIODIC, "Flux": _abi.OB_BC_FLUX, "Value": _abi.OB_BC_VALUE,
         "Gradient": _abi.OB_BC_GRADIENT, "Impenetrable": _abi.OB_BC_IMPENETRABLE, "Communication": _abi.OB_BC_COMMUNICATION}


def bc_desc(bcs):
    d = _abi.BcDesc()
    for s, side in enumerate(SIDES):
        bc = bcs[side]
        if bc is None:
            d.kind[s], d.value[s] = _abi.OB_BC_NONE, 0.0
        else:
            d.kind[s] = _KIND[bc.kind]
            d.value[s] = 0.0 if (bc.value is None or isinstance(bc.value, np.ndarray)) else float(bc.value)   # arrays: ob_model_set_bc_array
    return d


class Field:
    """Field{LX,LY,LZ}: `loc` is a 3-string of 'c'/'f'.  `data` is a device pointer to the parent array."""

    def __init__(self, grid, loc, bcs=None, name=""):
        self.grid, self.loc, self.name = grid, loc, name
        self.arch = grid.architecture
        self.n = tuple(grid.N[d] + (1 if (loc[d] == "f" and grid.topo[d] == _abi.OB_BOUNDED) else 0) for d in range(3))
        self.P = tuple(self.n[d] + 2 * grid.H[d] for d in range(3))
        self.count = self.P[0] * self.P[1] * self.P[2]
        self.nbytes = self.count * np.dtype(grid.FT).itemsize
        self.boundary_conditions = bcs if bcs is not None else regularize_bcs(grid, loc)
        p = C.c_void_p()
        _abi.call("ob_malloc", self.arch.ctx, self.nbytes, C.byref(p))  # zero-initialised
        self.data = p

    def __del__(self):
        try:
            if getattr(self, "data", None) and getattr(self.arch, "ctx", None):
                _abi.lib().ob_free(self.arch.ctx, self.data)
                self.data = None
        except Exception:
            pass

    # host <-> device (on_architecture(CPU(), parent(field)) / copyto!(parent, host)) -------------------------
    def parent(self):
        """host copy of the parent array, numpy shape (Pz, Py, Px) == column-major (Px, Py, Pz)"""
        out = np.empty(self.P[::-1], dtype=self.grid.FT)
        _abi.call("ob_memcpy_d2h", self.arch.ctx, out.ctypes.data_as(C.c_void_p), self.data, self.nbytes)
        return out

    def set_parent(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.grid.FT)
        if arr.shape != self.P[::-1]:
            raise ValueError("parent shape %r != %r" % (arr.shape, self.P[::-1]))
        _abi.call("ob_memcpy_h2d", self.arch.ctx, self.data, arr.ctypes.data_as(C.c_void_p), self.nbytes)
        _abi.call("ob_sync", self.arch.ctx)

    def interior(self):
        H = self.grid.H
        return self.parent()[H[2]:H[2] + self.n[2], H[1]:H[1] + self.n[1], H[0]:H[0] + self.n[0]]

    def nodes(self):
        g = self.grid
        return tuple(g.nodes(d, self.loc[d]) for d in range(3))

    def set(self, value):
        """set!(field, value): a number, an array of the interior size (numpy (nz,ny,nx) or Julia-ordered
        (nx,ny,nz)), or a function f(x, y, z) evaluated on the host nodes (set!.jl:81-125)."""
        g = self.grid
        H = g.H
        full = self.parent()
        inter = full[H[2]:H[2] + self.n[2], H[1]:H[1] + self.n[1], H[0]:H[0] + self.n[0]]
        if callable(value):
            x, y, z = self.nodes()
            X = x[None, None, :].astype(np.float64); Y = y[None, :, None].astype(np.float64); Z = z[:, None, None].astype(np.float64)
            nonflat = [a for a, t in zip((X, Y, Z), g.topo) if t != _abi.OB_FLAT]
            vals = np.vectorize(value, otypes=[np.float64])(*np.broadcast_arrays(*nonflat)) if nonflat else value()
            inter[...] = np.broadcast_to(vals, inter.shape)
        elif np.isscalar(value):
            inter[...] = value
        else:
            a = np.asarray(value)
            if a.shape == inter.shape:
                inter[...] = a
            elif a.shape == self.n:
                inter[...] = a.transpose(2, 1, 0)
            else:
                inter[...] = a.reshape(inter.shape)
        self.set_parent(full)

    def fill_halo_regions(self, fill_normal_flow_bcs=True):
        """fill_halo_regions!(field) for a field that is not (necessarily) bound to a model: ob_fill_halo_array.  Constant
        Flux / Value / Gradient conditions only (array-valued conditions go through the model: ob_model_set_bc_array)."""
        g = self.grid
        if getattr(self.arch, "world", 1) > 1:
            raise NotImplementedError("distributed halos are exchanged through the model (model.fill_halo_regions)")
        desc = g.desc()   # (keeps the host spacing arrays alive for the duration of the call)
        loc = (C.c_int32 * 3)(*[1 if c == "f" else 0 for c in self.loc])
        bcs = bc_desc(self.boundary_conditions)
        _abi.call("ob_fill_halo_array", self.arch.ctx, C.byref(desc), self.data, C.byref(loc), C.byref(bcs), None, int(bool(fill_normal_flow_bcs)))

    def any_nan(self):
        flag = C.c_int32(0)
        ft = _abi.OB_F64 if self.grid.FT == np.float64 else _abi.OB_F32
        _abi.call("ob_any_nan", self.arch.ctx, self.data, self.count, ft, C.byref(flag))
        from .distributed import all_reduce_scalar
        return bool(all_reduce_scalar(self.arch, flag.value, "max"))  # all-reduced across ranks (run.jl:191-197)

    def __repr__(self):
        return "%dx%dx%d Field{%s} %s" % (*self.n, self.loc, self.name)
