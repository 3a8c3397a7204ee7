"""ctypes binding of libjspace_b200.so (include/jspace_b200.h).

This is the Python twin of the Julia `ccall` wrapper in ../julia/JSpaceB200.jl: the same
symbols, the same structs.  The library is the only compute path — if it is missing the
import fails loudly (there is no Python or CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "JSPACE_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libjspace_b200.so"))

JSB_OK = 0
ERRORS = {-1: "JSB_EINVAL", -2: "JSB_ECUDA", -3: "JSB_ENOMEM", -4: "JSB_ECAPACITY",
          -5: "JSB_ENODEVICE", -6: "JSB_EIO"}

RULE_CONTACT, RULE_VOTER, RULE_HVOTER, RULE_NONE = 0, 1, 2, 3
METHOD_RANDOM, METHOD_TREE = 1, 2
QUIRK_EXCISION_INTENT = 0x1
QUIRK_DEBUG_SERIAL = 0x100
QUIRK_DEBUG_EXACT_CLASSES = 0x200
QUIRK_DEBUG_WIDE_GUARD = 0x400
QUIRK_DEBUG_VERIFY_TREE = 0x800
QUIRK_DEBUG_DECOUPLE = 0x1000
OUT_EVENTS, OUT_LATTICE, OUT_LISTS, OUT_SERIES = 0x1, 0x2, 0x4, 0x8
SERIES_MORE = 0x80000000
MODE_REJECTION_FREE = 0x100
EV_DUPLICATE, EV_MUTATION, EV_DEATH, EV_MIGRATION = 0, 1, 2, 3
EVF_GROUP_START, EVF_DISPLACED, EVF_EXCISION = 0x1, 0x2, 0x4
ST_OK, ST_NONTERMINATING, ST_CLONE_CAPACITY, ST_LOG_OVERFLOW, ST_MAX_ITER, ST_INTERNAL = range(6)

EVENT_DTYPE = np.dtype(
    [("time", "<f8"), ("cell", "<u4"), ("parent", "<u4"), ("site", "<u4"), ("site2", "<u4"),
     ("clone", "<u2"), ("label", "<u2"), ("type", "u1"), ("flags", "u1"), ("pad", "<u2")])
CLONE_DTYPE = np.dtype([("alpha", "<f8"), ("count", "<u4"), ("parent", "<u2"), ("label", "<u2")])
SUMMARY_DTYPE = np.dtype(
    [("n_iterations", "<u8"), ("n_events", "<u8"), ("t_final", "<f8"), ("n_effective", "<u4"),
     ("n_births", "<u4"), ("n_deaths", "<u4"), ("n_migrations", "<u4"), ("n_displaced", "<u4"),
     ("n_excised", "<u4"), ("n_alive", "<u4"), ("n_clones", "<u4"), ("next_cell_id", "<u4"),
     ("status", "<i4")])
SNAPSHOT_DTYPE = np.dtype([("time", "<f8"), ("iteration", "<u8"), ("n_events_before", "<u8"),
                           ("n_alive", "<u4"), ("index", "<u4")])
RELATION_DTYPE = np.dtype([("time", "<f8"), ("father", "<u4"), ("child", "<u4"),
                           ("subpop_child", "<u2"), ("type", "u1"), ("pad", "u1", (5,))])
SERIES_DTYPE = np.dtype([("n_alive", "<u4"), ("gain", "<u2"), ("loss", "<u2")])
assert EVENT_DTYPE.itemsize == 32 and CLONE_DTYPE.itemsize == 16 and SERIES_DTYPE.itemsize == 8
assert SUMMARY_DTYPE.itemsize == 64 and SNAPSHOT_DTYPE.itemsize == 32
assert RELATION_DTYPE.itemsize == 24


class Params(C.Structure):
    """jsb_params"""
    _fields_ = [
        ("size", C.c_uint32), ("method", C.c_uint32),
        ("Tf", C.c_double), ("rate_birth", C.c_double), ("rate_death", C.c_double),
        ("rate_migration", C.c_double), ("mu_driver", C.c_double), ("adv_mean", C.c_double),
        ("adv_std", C.c_double),
        ("rule", C.c_int32), ("n_start_cells", C.c_int32), ("n_bottleneck", C.c_int32),
        ("n_snapshots", C.c_int32),
        ("t_bottleneck", C.POINTER(C.c_double)), ("ratio_bottleneck", C.POINTER(C.c_double)),
        ("t_snapshot", C.POINTER(C.c_double)),
        ("n_tree_nodes", C.c_int32), ("n_tree_edges", C.c_int32),
        ("edge_parent", C.POINTER(C.c_int32)), ("edge_child", C.POINTER(C.c_int32)),
        ("node_rate", C.POINTER(C.c_double)),
        ("root_node", C.c_int32), ("quirk_flags", C.c_uint32),
        ("root_rate", C.c_double), ("max_iterations", C.c_uint64),
    ]


class RunOpts(C.Structure):
    """jsb_run_opts"""
    _fields_ = [
        ("size", C.c_uint32), ("flags", C.c_uint32), ("seed", C.c_uint64),
        ("g_begin", C.c_int64), ("g_end", C.c_int64),
        ("device", C.c_int32), ("warps_per_sm", C.c_int32), ("log_pool_bytes", C.c_uint64),
        ("n_devices", C.c_int32), ("reserved", C.c_int32), ("devices", C.POINTER(C.c_int32)),
    ]


# every symbol include/jspace_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_PVP = C.POINTER(C.c_void_p)
SIGNATURES = {
    "jsb_abi_version": (C.c_int, []),
    "jsb_device_count": (C.c_int, []),
    "jsb_last_error": (C.c_char_p, []),
    "jsb_graph_lattice": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _PVP]),
    "jsb_graph_csr": (C.c_int, [C.c_int64, _VP, _VP, _PVP]),
    "jsb_graph_dense_file": (C.c_int, [C.c_char_p, _PVP]),
    "jsb_graph_nv": (C.c_int64, [_VP]),
    "jsb_graph_ne": (C.c_int64, [_VP]),
    "jsb_graph_seed_candidates": (C.c_int, [_VP, _PVP, C.POINTER(C.c_int64)]),
    "jsb_graph_set_seed_candidates": (C.c_int, [_VP, _VP, C.c_int64]),
    "jsb_graph_neighbors": (C.c_int, [_VP, C.c_int64, _VP, C.c_int]),
    "jsb_graph_free": (None, [_VP]),
    "jsb_run": (C.c_int, [_VP, C.POINTER(Params), C.c_int64, C.c_int64, C.POINTER(RunOpts), _PVP]),
    "jsb_trim": (None, []),
    "jsb_result_kernel_ms": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "jsb_result_launches": (C.c_int, [_VP, C.POINTER(C.c_int64)]),
    "jsb_result_count": (C.c_int64, [_VP]),
    "jsb_result_summaries": (C.c_int, [_VP, _PVP]),
    "jsb_result_clones": (C.c_int, [_VP, C.c_int64, _PVP, C.POINTER(C.c_int32)]),
    "jsb_result_events": (C.c_int, [_VP, C.c_int64, _PVP, C.POINTER(C.c_int64)]),
    "jsb_result_lattice": (C.c_int, [_VP, C.c_int64, _PVP, _PVP]),
    "jsb_result_list": (C.c_int, [_VP, C.c_int64, C.c_int32, _PVP, C.POINTER(C.c_int64)]),
    "jsb_result_snapshots": (C.c_int, [_VP, C.c_int64, _PVP, C.POINTER(C.c_int32)]),
    "jsb_result_series": (C.c_int, [_VP, C.c_int64, _PVP, C.POINTER(C.c_int64), _PVP, C.POINTER(C.c_int32)]),
    "jsb_result_histogram": (C.c_int, [_VP, C.c_int64, C.c_int, _VP]),
    "jsb_result_free": (None, [_VP]),
    "jsb_result_genealogy": (C.c_int, [_VP, C.c_int64, _VP, C.c_int64, _PVP, C.POINTER(C.c_int64)]),
    "jsb_free": (None, [_VP]),
}


class JSpaceError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


_lib = None


def lib():
    """Loads libjspace_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C {os.path.dirname(_HERE)}` "
                "(or python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.jsb_abi_version() != 2:
            raise ImportError("libjspace_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != JSB_OK:
        raise JSpaceError(rc, lib().jsb_last_error().decode("utf-8", "replace"))


def view(ptr, count, dtype):
    """Copy `count` records at C pointer `ptr` into a numpy array."""
    dtype = np.dtype(dtype)
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer(C.string_at(ptr, int(count) * dtype.itemsize), dtype=dtype).copy()
