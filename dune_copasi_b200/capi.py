"""ctypes binding of include/dune_copasi_b200.h.

This is the stub a Python host would write against the C ABI (the reference's own host is C++; its
binding is shown in INTEGRATION.md).  Tests and bench.py go through these classes, i.e. through the
C ABI -- never around it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdune_copasi_b200.so")
_lib = None


class DcbError(RuntimeError):
    pass


class SolveResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("half_iterations", C.c_int32), ("converged", C.c_int32),
                ("reduction", C.c_double), ("defect0", C.c_double)]


class StepStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("steps", "failed_steps", "stages", "newton_iterations", "linear_solves",
                                        "linear_iterations", "linear_half_iterations", "residual_evaluations",
                                        "linearizations", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/dune_copasi_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)
SYMBOLS = {
    "dcb_version": (C.c_int, []),
    "dcb_last_error": (C.c_char_p, []),
    "dcb_device_count": (C.c_int, []),
    "dcb_config_create": (_P, []),
    "dcb_config_destroy": (None, [_P]),
    "dcb_config_parse_ini": (C.c_int, [_P, C.c_char_p]),
    "dcb_config_set": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "dcb_config_dump": (C.c_size_t, [_P, C.c_char_p, C.c_size_t]),
    "dcb_grid_create_structured": (_P, [C.c_int, _I32, _D, _D]),
    "dcb_grid_create_structured_cubes": (_P, [C.c_int, _I32, _D, _D]),
    "dcb_grid_nodes_per_element": (C.c_int, [_P]),
    "dcb_grid_create": (_P, [C.c_int, C.c_int64, _D, C.c_int64, _I32, C.c_int, C.POINTER(C.c_char_p), _D]),
    "dcb_grid_destroy": (None, [_P]),
    "dcb_grid_dim": (C.c_int, [_P]),
    "dcb_grid_num_vertices": (C.c_int64, [_P]),
    "dcb_grid_num_elements": (C.c_int64, [_P]),
    "dcb_grid_get_coords": (C.c_int, [_P, _D]),
    "dcb_grid_get_elements": (C.c_int, [_P, _I32]),
    "dcb_model_create": (_P, [_P, C.c_int, C.c_int, C.POINTER(C.c_char_p)]),
    "dcb_model_destroy": (None, [_P]),
    "dcb_model_num_compartments": (C.c_int, [_P]),
    "dcb_model_num_species": (C.c_int, [_P]),
    "dcb_model_species_name": (C.c_char_p, [_P, C.c_int]),
    "dcb_model_species_compartment": (C.c_int, [_P, C.c_int]),
    "dcb_model_cuda_source": (C.c_char_p, [_P]),
    "dcb_model_cuda_source_group": (C.c_char_p, [_P, C.c_int]),
    "dcb_operator_uses_tiles": (C.c_int, [_P]),
    "dcb_solver_is_fused": (C.c_int, [_P]),
    "dcb_model_compile": (C.c_int64, [_P, C.c_int, C.c_char_p, C.c_size_t]),
    "dcb_model_precompile": (C.c_int, [_P]),
    "dcb_model_precompile_group": (C.c_int, [_P, C.c_int]),
    "dcb_grid_bind": (C.c_int, [_P, _P]),
    "dcb_grid_num_dofs": (C.c_int64, [_P]),
    "dcb_grid_get_elem_compartment": (C.c_int, [_P, _I32]),
    "dcb_grid_get_elem_dof": (C.c_int, [_P, _I64]),
    "dcb_grid_num_facets": (C.c_int64, [_P]),
    "dcb_grid_get_facets": (C.c_int, [_P, _I64, _I64, _I32, _I32]),
    "dcb_grid_pattern": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32]),
    "dcb_grid_interpolate": (C.c_int, [_P, _P, C.c_double, _D]),
    "dcb_grid_write_vtk": (C.c_int, [_P, _P, _D, C.c_double, C.c_char_p, C.c_int]),
    "dcb_grid_constraints": (C.c_int64, [_P, _P, _I32, _D, C.c_int64]),
    "dcb_operator_create": (_P, [_P, _P]),
    "dcb_operator_destroy": (None, [_P]),
    "dcb_operator_num_dofs": (C.c_int64, [_P]),
    "dcb_operator_nnz": (C.c_int64, [_P]),
    "dcb_operator_launches": (C.c_int64, [_P]),
    "dcb_operator_stream": (_P, [_P]),
    "dcb_operator_sync": (C.c_int, [_P]),
    "dcb_operator_profile": (C.c_int, [_P, C.c_int]),
    "dcb_operator_profile_report": (C.c_size_t, [_P, C.c_char_p, C.c_size_t]),
    "dcb_residual": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _D, _D]),
    "dcb_jacobian": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _D, _D]),
    "dcb_jacobian_apply": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _D, _D, _D]),
    "dcb_block_diagonal": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _D, _D, C.c_int64]),
    "dcb_residual_dev": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _P, _P]),
    "dcb_jacobian_dev": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _P, _P]),
    "dcb_jacobian_apply_dev": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "dcb_solver_create": (_P, [_P, _P, _P]),
    "dcb_solver_destroy": (None, [_P]),
    "dcb_solver_linearize": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _D]),
    "dcb_solver_solve": (C.c_int, [_P, _D, _D, C.c_double, C.POINTER(SolveResult)]),
    "dcb_solver_apply_operator": (C.c_int, [_P, _D, _D]),
    "dcb_stepper_create": (_P, [_P, _P, _P]),
    "dcb_stepper_destroy": (None, [_P]),
    "dcb_stepper_set_state": (C.c_int, [_P, _D, C.c_double]),
    "dcb_stepper_get_state": (C.c_int, [_P, _D, _D]),
    "dcb_stepper_state_dev": (_P, [_P]),
    "dcb_stepper_set_time": (C.c_int, [_P, C.c_double]),
    "dcb_stepper_step": (C.c_int, [_P, C.c_double, C.POINTER(C.c_int)]),
    "dcb_stepper_evolve": (C.c_int, [_P, C.c_double, _D, C.c_int, C.POINTER(C.c_int)]),
    "dcb_stepper_stats": (C.c_int, [_P, C.POINTER(StepStats)]),
    "dcb_reducer_create": (_P, [_P, _P, _P]),
    "dcb_reducer_destroy": (None, [_P]),
    "dcb_reducer_num_keys": (C.c_int, [_P]),
    "dcb_reducer_key": (C.c_char_p, [_P, C.c_int]),
    "dcb_reducer_apply": (C.c_int, [_P, C.c_double, _D, _D, _I32]),
    "dcb_reducer_apply_dev": (C.c_int, [_P, C.c_double, _P, _D, _I32]),
    "dcb_model_precompile_reduce": (C.c_int, [_P, _P]),
    "dcb_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "dcb_grid_partition": (_P, [_P, C.c_int, C.c_int]),
    "dcb_grid_partition_method": (_P, [_P, C.c_int, C.c_int, C.c_char_p]),
    "dcb_grid_num_owned_vertices": (C.c_int64, [_P]),
    "dcb_grid_owned_vertex_range": (C.c_int, [_P, _I64, _I64]),
    "dcb_grid_get_global_vertex_ids": (C.c_int, [_P, _I64]),
    "dcb_grid_get_vertex_owner": (C.c_int, [_P, _I32]),
    "dcb_grid_get_global_element_ids": (C.c_int, [_P, _I64]),
    "dcb_grid_halo_num_peers": (C.c_int, [_P, C.c_int]),
    "dcb_grid_halo_peer": (C.c_int, [_P, C.c_int, C.c_int, _I32, _I64, _I64]),
    "dcb_grid_halo_lists": (C.c_int, [_P, C.c_int, C.c_int, _I32, _I32]),
    "dcb_comm_create": (_P, [C.c_char_p, C.c_int, C.c_int, _P]),
    "dcb_comm_destroy": (None, [_P]),
    "dcb_comm_uses_peer_memory": (C.c_int, [_P]),
    "dcb_operator_owned_ranges": (C.c_int, [_P, _I64, _I64, C.c_int]),
}


def lib():
    """Load the in-tree library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise DcbError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise DcbError(lib().dcb_last_error().decode())


def _ptr(obj, what):
    if not obj:
        raise DcbError(f"{what}: {lib().dcb_last_error().decode()}")
    return obj


def _d(a):
    return a.ctypes.data_as(_D)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Config:
    def __init__(self, ini_text: str = "", **overrides):
        self.h = _ptr(lib().dcb_config_create(), "config")
        if ini_text:
            _check(lib().dcb_config_parse_ini(self.h, ini_text.encode()))
        for k, v in overrides.items():
            self.set(k, v)

    def set(self, key, value):
        _check(lib().dcb_config_set(self.h, key.encode(), str(value).encode()))
        return self

    def dump(self) -> str:
        n = lib().dcb_config_dump(self.h, None, 0)
        buf = C.create_string_buffer(n)
        lib().dcb_config_dump(self.h, buf, n)
        return buf.value.decode()

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_config_destroy(self.h)


def _keys(keys):
    arr = (C.c_char_p * max(1, len(keys)))(*[k.encode() for k in keys])
    return arr


class Grid:
    def __init__(self, handle):
        self.h = _ptr(handle, "grid")

    @staticmethod
    def structured(dim, cells, origin=None, extent=None, element="simplex"):
        """element = "simplex": Kuhn split of the lattice (the reference's grid); "cube": the cells
        as Q1 elements (BASELINE configs[3]; not a reference capability)."""
        cells = np.asarray(cells, dtype=np.int32)
        origin = _f64(np.zeros(dim) if origin is None else origin)
        extent = _f64(np.ones(dim) if extent is None else extent)
        if element not in ("simplex", "cube"):
            raise ValueError("element must be 'simplex' or 'cube'")
        make = lib().dcb_grid_create_structured_cubes if element == "cube" else lib().dcb_grid_create_structured
        return Grid(make(dim, cells.ctypes.data_as(_I32), _d(origin), _d(extent)))

    @staticmethod
    def from_arrays(dim, coords, elems, cell_keys=(), cell_data=None):
        coords = _f64(coords)
        elems = np.ascontiguousarray(elems, dtype=np.int32)
        cd = _f64(cell_data) if cell_data is not None and len(cell_keys) else None
        return Grid(lib().dcb_grid_create(dim, coords.shape[0], _d(coords), elems.shape[0],
                                          elems.ctypes.data_as(_I32), len(cell_keys), _keys(list(cell_keys)),
                                          _d(cd) if cd is not None else None))

    dim = property(lambda s: lib().dcb_grid_dim(s.h))
    nv = property(lambda s: lib().dcb_grid_num_vertices(s.h))
    ne = property(lambda s: lib().dcb_grid_num_elements(s.h))
    ndofs = property(lambda s: lib().dcb_grid_num_dofs(s.h))
    nodes_per_element = property(lambda s: lib().dcb_grid_nodes_per_element(s.h))

    def coords(self):
        out = np.empty((self.nv, self.dim))
        lib().dcb_grid_get_coords(self.h, _d(out))
        return out

    def elements(self):
        out = np.empty((self.ne, self.nodes_per_element), dtype=np.int32)
        lib().dcb_grid_get_elements(self.h, out.ctypes.data_as(_I32))
        return out

    def bind(self, model):
        _check(lib().dcb_grid_bind(self.h, model.h))
        return self

    def elem_compartment(self):
        out = np.empty(self.ne, dtype=np.int32)
        lib().dcb_grid_get_elem_compartment(self.h, out.ctypes.data_as(_I32))
        return out

    def elem_dof(self):
        out = np.empty((self.ne, self.nodes_per_element), dtype=np.int64)
        lib().dcb_grid_get_elem_dof(self.h, out.ctypes.data_as(_I64))
        return out

    def facets(self):
        n = lib().dcb_grid_num_facets(self.h)
        fi, fo = np.empty(n, np.int64), np.empty(n, np.int64)
        li, lo = np.empty(n, np.int32), np.empty(n, np.int32)
        if n:
            lib().dcb_grid_get_facets(self.h, fi.ctypes.data_as(_I64), fo.ctypes.data_as(_I64),
                                      li.ctypes.data_as(_I32), lo.ctypes.data_as(_I32))
        return fi, fo, li, lo

    def pattern(self, model):
        nrows, nnz = C.c_int64(), C.c_int64()
        _check(lib().dcb_grid_pattern(self.h, model.h, C.byref(nrows), C.byref(nnz), None, None))
        rp = np.empty(nrows.value + 1, dtype=np.int64)
        ci = np.empty(nnz.value, dtype=np.int32)
        _check(lib().dcb_grid_pattern(self.h, model.h, None, None, rp.ctypes.data_as(_I64), ci.ctypes.data_as(_I32)))
        return rp, ci

    def interpolate(self, model, time):
        u = np.empty(self.ndofs)
        _check(lib().dcb_grid_interpolate(self.h, model.h, time, _d(u)))
        return u

    def write_vtk(self, model, u, time, path, append=True):
        u = _f64(u)
        _check(lib().dcb_grid_write_vtk(self.h, model.h, _d(u), time, str(path).encode(), 1 if append else 0))

    def constraints(self, model):
        n = lib().dcb_grid_constraints(self.h, model.h, None, None, 0)
        if n < 0:
            raise DcbError(lib().dcb_last_error().decode())
        d, v = np.empty(n, np.int32), np.empty(n)
        if n:
            lib().dcb_grid_constraints(self.h, model.h, d.ctypes.data_as(_I32), _d(v), n)
        return d, v

    def partition(self, rank, size, method="auto"):
        """slab (structured lattices) | rcb | range | auto (slab for lattices, rcb otherwise)"""
        return Grid(lib().dcb_grid_partition_method(self.h, rank, size, method.encode()))

    def global_element_ids(self):
        out = np.empty(self.ne, dtype=np.int64)
        lib().dcb_grid_get_global_element_ids(self.h, out.ctypes.data_as(_I64))
        return out

    def halo_plan(self, rank):
        """-> [(peer, send dof indices, recv dof indices)]"""
        out = []
        for k in range(lib().dcb_grid_halo_num_peers(self.h, rank)):
            peer, ns, nr = C.c_int32(), C.c_int64(), C.c_int64()
            lib().dcb_grid_halo_peer(self.h, rank, k, C.byref(peer), C.byref(ns), C.byref(nr))
            s, r = np.empty(ns.value, np.int32), np.empty(nr.value, np.int32)
            lib().dcb_grid_halo_lists(self.h, rank, k, s.ctypes.data_as(_I32), r.ctypes.data_as(_I32))
            out.append((peer.value, s, r))
        return out

    n_owned = property(lambda s: lib().dcb_grid_num_owned_vertices(s.h))

    def owned_vertex_range(self):
        b, e = C.c_int64(), C.c_int64()
        lib().dcb_grid_owned_vertex_range(self.h, C.byref(b), C.byref(e))
        return b.value, e.value

    def global_vertex_ids(self):
        out = np.empty(self.nv, dtype=np.int64)
        lib().dcb_grid_get_global_vertex_ids(self.h, out.ctypes.data_as(_I64))
        return out

    def vertex_owner(self):
        out = np.empty(self.nv, dtype=np.int32)
        lib().dcb_grid_get_vertex_owner(self.h, out.ctypes.data_as(_I32))
        return out

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_grid_destroy(self.h)


class Model:
    def __init__(self, config: Config, dim: int, cell_keys=()):
        self.config = config
        self.h = _ptr(lib().dcb_model_create(config.h, dim, len(cell_keys), _keys(list(cell_keys))), "model")

    nspec = property(lambda s: lib().dcb_model_num_species(s.h))
    ncomp = property(lambda s: lib().dcb_model_num_compartments(s.h))

    def species(self):
        return [(lib().dcb_model_species_name(self.h, i).decode(), lib().dcb_model_species_compartment(self.h, i))
                for i in range(self.nspec)]

    def cuda_source(self) -> str:
        s = lib().dcb_model_cuda_source(self.h)
        if s is None:
            raise DcbError(lib().dcb_last_error().decode())
        return s.decode()

    def precompile_group(self, group: int):
        _check(lib().dcb_model_precompile_group(self.h, group))

    def cuda_source_group(self, group: int) -> str:
        s = lib().dcb_model_cuda_source_group(self.h, group)
        if s is None:
            _check(-1)
        return s.decode()

    def precompile(self, reduce_config: "Config | None" = None):
        _check(lib().dcb_model_precompile(self.h))
        if reduce_config is not None:
            _check(lib().dcb_model_precompile_reduce(self.h, reduce_config.h))

    def compile(self, ptx=False) -> bytes:
        n = lib().dcb_model_compile(self.h, 1 if ptx else 0, None, 0)
        if n < 0:
            raise DcbError(lib().dcb_last_error().decode())
        buf = C.create_string_buffer(n)
        lib().dcb_model_compile(self.h, 1 if ptx else 0, buf, n)
        return buf.raw

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_model_destroy(self.h)


class Operator:
    """Device operator; host-array methods copy inside the call (the reference-facing path)."""

    def __init__(self, model: Model, grid: Grid):
        self.model, self.grid = model, grid
        self.h = _ptr(lib().dcb_operator_create(model.h, grid.h), "operator")

    ndofs = property(lambda s: lib().dcb_operator_num_dofs(s.h))
    nnz = property(lambda s: lib().dcb_operator_nnz(s.h))
    launches = property(lambda s: lib().dcb_operator_launches(s.h))
    stream = property(lambda s: lib().dcb_operator_stream(s.h))

    @property
    def uses_tiles(self) -> bool:
        return bool(lib().dcb_operator_uses_tiles(self.h))

    def sync(self):
        _check(lib().dcb_operator_sync(self.h))

    def residual(self, time, wM, wA, x, r=None):
        x = _f64(x)
        r = np.zeros(self.ndofs) if r is None else _f64(r)
        _check(lib().dcb_residual(self.h, time, wM, wA, _d(x), _d(r)))
        return r

    def jacobian(self, time, wM, wA, x):
        x = _f64(x)
        vals = np.empty(self.nnz)
        _check(lib().dcb_jacobian(self.h, time, wM, wA, _d(x), _d(vals)))
        return vals

    def jacobian_apply(self, time, wM, wA, x, z, y=None):
        x, z = _f64(x), _f64(z)
        y = np.zeros(self.ndofs) if y is None else _f64(y)
        _check(lib().dcb_jacobian_apply(self.h, time, wM, wA, _d(x), _d(z), _d(y)))
        return y

    def block_diagonal(self, time, wM, wA, x, size):
        x = _f64(x)
        out = np.empty(size)
        _check(lib().dcb_block_diagonal(self.h, time, wM, wA, _d(x), _d(out), size))
        return out

    def residual_dev(self, time, wM, wA, x_ptr, r_ptr):
        _check(lib().dcb_residual_dev(self.h, time, wM, wA, x_ptr, r_ptr))

    def jacobian_apply_dev(self, time, wM, wA, x_ptr, z_ptr, y_ptr):
        _check(lib().dcb_jacobian_apply_dev(self.h, time, wM, wA, x_ptr, z_ptr, y_ptr))

    def jacobian_dev(self, time, wM, wA, x_ptr, vals_ptr):
        _check(lib().dcb_jacobian_dev(self.h, time, wM, wA, x_ptr, vals_ptr))

    def profile(self):
        """-> {kind: {"ms": accumulated device ms, "launches": n}} since profiling was enabled"""
        buf = C.create_string_buffer(1 << 16)
        lib().dcb_operator_profile_report(self.h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            k, ms, n = line.split()
            out[k] = {"ms": float(ms), "launches": int(n)}
        return out

    def owned_ranges(self):
        b, e = np.zeros(8, np.int64), np.zeros(8, np.int64)
        n = lib().dcb_operator_owned_ranges(self.h, b.ctypes.data_as(_I64), e.ctypes.data_as(_I64), 8)
        return list(zip(b[:n].tolist(), e[:n].tolist()))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_operator_destroy(self.h)


class Comm:
    def __init__(self, unique_id: bytes, rank: int, size: int, op: Operator):
        self.h = _ptr(lib().dcb_comm_create(unique_id, rank, size, op.h), "comm")

    @property
    def uses_peer_memory(self) -> bool:
        return bool(lib().dcb_comm_uses_peer_memory(self.h))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().dcb_nccl_unique_id(buf))
        return buf.raw

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_comm_destroy(self.h)


class Solver:
    def __init__(self, op: Operator, linear_solver_cfg: Config, comm: Comm | None = None):
        self.op = op
        self.h = _ptr(lib().dcb_solver_create(op.h, linear_solver_cfg.h, comm.h if comm else None), "solver")

    @property
    def fused(self) -> bool:
        return bool(lib().dcb_solver_is_fused(self.h))

    def linearize(self, time, wM, wA, x):
        x = _f64(x)
        _check(lib().dcb_solver_linearize(self.h, time, wM, wA, _d(x)))

    def solve(self, b, rel_tol):
        b = _f64(b)
        z = np.empty(self.op.ndofs)
        res = SolveResult()
        _check(lib().dcb_solver_solve(self.h, _d(b), _d(z), rel_tol, C.byref(res)))
        return z, res

    def apply_operator(self, v):
        v = _f64(v)
        y = np.empty(self.op.ndofs)
        _check(lib().dcb_solver_apply_operator(self.h, _d(v), _d(y)))
        return y

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_solver_destroy(self.h)


class Stepper:
    def __init__(self, op: Operator, config: Config, comm: Comm | None = None):
        self.op = op
        self.h = _ptr(lib().dcb_stepper_create(op.h, config.h, comm.h if comm else None), "stepper")

    def set_state(self, u, time):
        u = _f64(u)
        _check(lib().dcb_stepper_set_state(self.h, _d(u), time))

    def get_state(self, out=None):
        u = np.empty(self.op.ndofs) if out is None else out
        t = C.c_double()
        _check(lib().dcb_stepper_get_state(self.h, _d(u), C.byref(t)))
        return u, t.value

    @property
    def time(self):
        t = C.c_double()
        _check(lib().dcb_stepper_get_state(self.h, None, C.byref(t)))
        return t.value

    def state_dev(self):
        return lib().dcb_stepper_state_dev(self.h)

    def step(self, dt) -> bool:
        ok = C.c_int()
        _check(lib().dcb_stepper_step(self.h, dt, C.byref(ok)))
        return bool(ok.value)

    def evolve(self, t_end, dt, max_steps=1 << 30):
        dtc, acc = C.c_double(dt), C.c_int()
        _check(lib().dcb_stepper_evolve(self.h, t_end, C.byref(dtc), max_steps, C.byref(acc)))
        return acc.value, dtc.value

    def stats(self):
        s = StepStats()
        lib().dcb_stepper_stats(self.h, C.byref(s))
        return s.as_dict()

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_stepper_destroy(self.h)


class ReductionError(DcbError):
    """An error.expression of [model.reduce] fired (the reference's ReductionError, reduce.hh:36)."""


class Reducer:
    """[model.reduce] functionals; `apply` returns {key: value} and raises ReductionError like the
    reference when an error expression fires (values and statuses stay readable in .values/.status)."""

    def __init__(self, op: Operator, config: Config, comm: Comm | None = None):
        self.op = op
        self.h = _ptr(lib().dcb_reducer_create(op.h, config.h, comm.h if comm else None), "reducer")
        n = lib().dcb_reducer_num_keys(self.h)
        self.keys = [lib().dcb_reducer_key(self.h, k).decode() for k in range(n)]
        self.values, self.status = {}, {}

    def _finish(self, rc, vals, stat, raise_on_error):
        if rc == 1:
            raise DcbError(lib().dcb_last_error().decode())
        self.values = dict(zip(self.keys, vals.tolist()))
        self.status = dict(zip(self.keys, stat.tolist()))
        if rc == 2 and raise_on_error:
            raise ReductionError(lib().dcb_last_error().decode())
        return self.values

    def apply(self, time, x, raise_on_error=True):
        x = _f64(x)
        vals, stat = np.zeros(len(self.keys)), np.zeros(len(self.keys), dtype=np.int32)
        rc = lib().dcb_reducer_apply(self.h, time, _d(x), _d(vals), stat.ctypes.data_as(_I32))
        return self._finish(rc, vals, stat, raise_on_error)

    def apply_dev(self, time, x_dev, raise_on_error=True):
        vals, stat = np.zeros(len(self.keys)), np.zeros(len(self.keys), dtype=np.int32)
        rc = lib().dcb_reducer_apply_dev(self.h, time, x_dev, _d(vals), stat.ctypes.data_as(_I32))
        return self._finish(rc, vals, stat, raise_on_error)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.dcb_reducer_destroy(self.h)
