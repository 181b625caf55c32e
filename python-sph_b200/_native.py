"""ctypes binding of libsphb200.so (the C-ABI declared in include/sphb200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, every
compute call raises.  The library is built in-tree by ``__graft_entry__.build()`` /
``python python-sph_b200/build_native.py``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPHB_LIB: a variant library built by `build_native.py --out ... -D...` (tuning sweeps); never set in tests or bench
LIB_PATH = os.environ.get("SPHB_LIB") or os.path.join(_HERE, "libsphb200.so")

# status bits (include/sphb200.h SPHB_ST_*)
ST_NEG_H, ST_NEG_P, ST_NEG_PI, ST_NEG_VISC, ST_NEG_E, ST_BISECT_NONE = 1, 2, 4, 8, 16, 32


class Consts(C.Structure):
    """sphb_consts: pyfluid.py:5-41"""
    _fields_ = [
        ("particle_mass", C.c_double), ("coupling_const", C.c_double), ("h_tol", C.c_double),
        ("newton_limit", C.c_int32), ("bisect_limit", C.c_int32),
        ("alpha", C.c_double), ("beta", C.c_double), ("epsilon", C.c_double),
        ("bisect_tol", C.c_double), ("bisect_a", C.c_double), ("bisect_b", C.c_double),
        ("G", C.c_double), ("gamma", C.c_double), ("mass", C.c_void_p),
    ]


class Grid(C.Structure):
    _fields_ = [("ox", C.c_double), ("oy", C.c_double), ("oz", C.c_double), ("cell", C.c_double),
                ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("xdiv", C.c_int32)]

    @property
    def ncell(self):
        return int(self.nx) * int(self.ny) * int(self.nz)


class Status(C.Structure):
    _fields_ = [("flags", C.c_int32), ("newton_warn", C.c_int32), ("bisect_oob", C.c_int32),
                ("bisect_none", C.c_int32), ("nfail", C.c_int32), ("reserved", C.c_int32 * 3)]


STATUS_INTS = C.sizeof(Status) // 4


class Level(C.Structure):
    """sphb_level: one of the nested grids (cells numbered from cell_base)"""
    _fields_ = [("grid", Grid), ("cell_base", C.c_uint32), ("ncell", C.c_uint32)]


MAX_LEVELS = 16


class Cells(C.Structure):
    _fields_ = [("grid", Grid), ("n", C.c_int64), ("cell_start", C.c_void_p), ("cell_hmax", C.c_void_p),
                ("cell_id", C.c_void_p), ("posh", C.c_void_p), ("frec", C.c_void_p), ("delta", C.c_float),
                ("pad_", C.c_float), ("nlevels", C.c_int32), ("exp0", C.c_int32), ("levels", C.c_void_p),
                ("level_hmax", C.c_void_p)]


class Lists(C.Structure):
    """sphb_lists: neighbour lists parked between two kernels on one snapshot"""
    _fields_ = [("store", C.c_void_p), ("broken", C.c_void_p), ("hmax_global", C.c_void_p), ("cover", C.c_double)]


class NList(C.Structure):
    """sphb_nlist: neighbour lists kept across kernels and steps"""
    _fields_ = [("n", C.c_int64), ("warp_off", C.c_void_p), ("warp_rows", C.c_void_p), ("pool", C.c_void_p),
                ("pool_rows", C.c_int64), ("cursor", C.c_void_p), ("flags", C.c_void_p), ("posh_build", C.c_void_p),
                ("cover_h", C.c_double), ("skin", C.c_double), ("maxima", C.c_void_p), ("margins", C.c_void_p),
                ("level_dmax", C.c_void_p), ("cell_dmax", C.c_void_p)]


NL_NONE = 0xFFFFFFFF
NL_POOL, NL_BROKEN, NL_AGING, NL_UNLISTABLE = 1, 2, 4, 8

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32

# name -> (restype, argtypes); every symbol of include/sphb200.h
SIGNATURES = {
    "sphb_last_cuda_error": (C.c_char_p, []),
    "sphb_abi_version": (C.c_int, []),
    "sphb_device_ok": (C.c_int, []),
    "sphb_status_reset": (C.c_int, [P, P]),
    "sphb_grid_ncell": (I64, [C.POINTER(Grid)]),
    "sphb_grid_workspace_bytes": (I64, [I64, C.POINTER(Grid)]),
    "sphb_pack_posh": (C.c_int, [P, P, I64, P, P]),
    "sphb_grid_build": (C.c_int, [P, P, I64, C.POINTER(Grid), C.c_float, P, P, P, P, P, P, P, I64, P]),
    "sphb_grid_set_h": (C.c_int, [P, I64, C.POINTER(Grid), C.c_float, P, P, P, P, P]),
    "sphb_levels_workspace_bytes": (I64, [I64, I64]),
    "sphb_grid_build_levels": (C.c_int, [P, P, I64, C.POINTER(Cells), I64, P, P, I64, P]),
    "sphb_level_stats": (C.c_int, [P, I64, I32, P, P]),
    "sphb_grid_set_h_levels": (C.c_int, [P, C.POINTER(Cells), I64, P]),
    "sphb_gather_f64": (C.c_int, [P, P, I64, I32, P, P]),
    "sphb_scatter_f64": (C.c_int, [P, P, I64, I32, P, P]),
    "sphb_gather_u32": (C.c_int, [P, P, I64, P, P]),
    "sphb_lists_bytes": (C.c_size_t, [I64]),
    "sphb_nlist_pool_rows": (I64, [I64]),
    "sphb_sizeof_nlist": (I64, []),
    "sphb_nlist_build": (C.c_int, [C.POINTER(Cells), P, P, C.POINTER(NList), C.POINTER(Consts), P, P]),
    "sphb_nlist_check": (C.c_int, [C.POINTER(NList), P, P]),
    "sphb_nlist_mark": (C.c_int, [C.POINTER(NList), P, P]),
    "sphb_nlist_remap": (C.c_int, [C.POINTER(NList), P, P]),
    "sphb_nlist_density_omega": (C.c_int, [C.POINTER(NList), C.POINTER(Consts), P, P, P, P, P, P, P, P]),
    "sphb_nlist_newton_h": (C.c_int, [C.POINTER(NList), C.POINTER(Consts), P, P, P, P, P, P, P, P, P]),
    "sphb_nlist_eos": (C.c_int, [P, P, P, P, I64, C.POINTER(Consts), P, P, P, P]),
    "sphb_nlist_force": (C.c_int, [C.POINTER(NList), C.POINTER(Consts), P, P, P, P, P, P, P, P, P]),
    "sphb_rows_pack": (C.c_int, [P, I32, P, I32, P, I32, P, I32, P, I64, P, I32, P]),
    "sphb_gather_state": (C.c_int, [P, P, P, P, P, I64, P, P, P, P]),
    "sphb_scatter_state": (C.c_int, [P, P, P, P, I64, P, P, P, P, P]),
    "sphb_density_omega": (C.c_int, [C.POINTER(Cells), C.POINTER(Consts), P, P, P, P, P, P, P, P, C.POINTER(Lists), P]),
    "sphb_newton_h": (C.c_int, [C.POINTER(Cells), C.POINTER(Consts), P, P, P, P, P, P, P, P, C.POINTER(Lists), P]),
    "sphb_bisect_h": (C.c_int, [C.POINTER(Cells), C.POINTER(Consts), P, I64, P, P, P, P, I32, P, P]),
    "sphb_neighbours": (C.c_int, [C.POINTER(Cells), P, I64, I32, P, P, I64, P, P]),
    "sphb_eos": (C.c_int, [P, P, P, P, P, P, P, I64, C.POINTER(Consts), P, P, P, P, P, P, P]),
    "sphb_force": (C.c_int, [C.POINTER(Cells), C.POINTER(Consts), P, P, P, P, P, P, P, P, C.POINTER(Lists), P]),
    "sphb_gravity_direct": (C.c_int, [P, I64, C.POINTER(Consts), I32, P, P]),
    "sphb_gravity_direct_targets": (C.c_int, [P, I64, P, I64, C.POINTER(Consts), I32, P, P]),
    "sphb_grav_eval": (C.c_int, [P, P, I64, P, P, P, P]),
    "sphb_pair_terms": (C.c_int, [P, P, P, P, I64, C.POINTER(Consts), P, P, P, P, P, P, P]),
    "sphb_predict": (C.c_int, [P, P, P, P, P, C.c_double, I64, P, P, P, P, P]),
    "sphb_correct": (C.c_int, [P, P, P, P, P, C.c_double, I64, P, P, P, P, P]),
    "sphb_bounds": (C.c_int, [P, I64, P, P, I64, P]),
    "sphb_bounds_workspace_bytes": (I64, [I64]),
    "sphb_hmax": (C.c_int, [P, I64, P, P]),
    "sphb_distance": (C.c_int, [P, P, I64, P, P]),
    "sphb_m4_eval": (C.c_int, [P, P, I64, P, P, P, P]),
    "sphb_check_neg_h": (C.c_int, [P, I64, P, P]),
    "sphb_fp64_peak": (C.c_int, [I32, I32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sphb_launch_count": (I64, []),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """dlopen libsphb200.so and set the prototypes.  Raises NativeError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root); "
            "sph-b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sphb_abi_version() != 1:
        raise NativeError("libsphb200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc == 0:
        return
    msg = {-1: "bad argument", -2: "CUDA error: " + (load().sphb_last_cuda_error() or b"").decode(), -3: "no sm_100 device"}
    raise NativeError(f"libsphb200 {what}: {msg.get(rc, rc)}")
