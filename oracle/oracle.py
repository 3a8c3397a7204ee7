"""ctypes access to the CPU oracle (oracle/jspace_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libjspace_oracle.so")

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
assert EVENT_DTYPE.itemsize == 32 and CLONE_DTYPE.itemsize == 16
assert SUMMARY_DTYPE.itemsize == 64 and SNAPSHOT_DTYPE.itemsize == 32


class OrcParams(C.Structure):
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


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("jspace_oracle.cpp", "julia_mt.hpp", "zig_tables.inc")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_log.restype = C.c_double
        L.orc_log.argtypes = [C.c_double]
        L.orc_graph_lattice.restype = C.c_void_p
        L.orc_graph_lattice.argtypes = [C.c_int, C.c_int64, C.c_int64]
        L.orc_graph_csr.restype = C.c_void_p
        L.orc_graph_csr.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_graph_dense_file.restype = C.c_void_p
        L.orc_graph_dense_file.argtypes = [C.c_char_p]
        L.orc_graph_nv.restype = C.c_int64
        L.orc_graph_nv.argtypes = [C.c_void_p]
        L.orc_graph_neighbors.restype = C.c_int
        L.orc_graph_neighbors.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.orc_graph_closeness_argmax.restype = C.c_int64
        L.orc_graph_closeness_argmax.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_run.restype = C.c_void_p
        L.orc_run.argtypes = [C.c_void_p, C.POINTER(OrcParams), C.c_uint64, C.c_uint32, C.c_void_p,
                              C.c_int64, C.c_int]
        L.orc_summary.restype = C.c_void_p
        L.orc_summary.argtypes = [C.c_void_p]
        for name in ("orc_events", "orc_clones", "orc_snapshots", "orc_series"):
            f = getattr(L, name)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.orc_lattice.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_list.restype = C.c_int64
        L.orc_list.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def contract_log(x: float) -> float:
    return lib().orc_log(float(x))


RULES = {"contact": 0, "voter": 1, "h_voter": 2, "hvoter": 2}


def make_params(*, method=1, Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01,
                mu_driver=0.01, adv_mean=0.4, adv_std=0.1, rule="contact", n_start_cells=1,
                t_bottleneck=(), ratio_bottleneck=(), t_snapshot=(), tree_edges=(), node_rate=(),
                root_node=0, root_rate=0.0, quirk_flags=0, max_iterations=0):
    """Returns (OrcParams, keepalive) — keepalive holds the numpy buffers the struct points to."""
    p = OrcParams()
    p.size = C.sizeof(OrcParams)
    p.method = method
    p.Tf, p.rate_birth, p.rate_death, p.rate_migration = Tf, rate_birth, rate_death, rate_migration
    p.mu_driver, p.adv_mean, p.adv_std = mu_driver, adv_mean, adv_std
    p.rule = RULES.get(rule, 3) if isinstance(rule, str) else int(rule)
    p.n_start_cells = n_start_cells
    keep = []

    def darr(v):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double)) if a.size else None

    def iarr(v):
        a = np.ascontiguousarray(np.asarray(v, dtype=np.int32))
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32)) if a.size else None

    p.n_bottleneck = len(t_bottleneck)
    p.t_bottleneck = darr(t_bottleneck)
    p.ratio_bottleneck = darr(ratio_bottleneck)
    p.n_snapshots = len(t_snapshot)
    p.t_snapshot = darr(t_snapshot)
    p.n_tree_nodes = len(node_rate)
    p.n_tree_edges = len(tree_edges)
    p.edge_parent = iarr([e[0] for e in tree_edges])
    p.edge_child = iarr([e[1] for e in tree_edges])
    p.node_rate = darr(node_rate)
    p.root_node = root_node
    p.root_rate = root_rate
    p.quirk_flags = quirk_flags
    p.max_iterations = max_iterations
    return p, keep


class Graph:
    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle graph construction failed")
        self.h = handle
        self.nv = lib().orc_graph_nv(handle)

    @classmethod
    def lattice(cls, dim, row, col):
        return cls(lib().orc_graph_lattice(dim, row, col))

    @classmethod
    def csr(cls, rowptr, colidx):
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(colidx, dtype=np.int32)
        return cls(lib().orc_graph_csr(len(rp) - 1, rp.ctypes.data, ci.ctypes.data))

    @classmethod
    def dense_file(cls, path):
        return cls(lib().orc_graph_dense_file(os.fsencode(path)))

    def neighbors(self, v):
        out = np.zeros(64, dtype=np.int64)
        d = lib().orc_graph_neighbors(self.h, v, out.ctypes.data, 64)
        if d > 64:
            out = np.zeros(d, dtype=np.int64)
            lib().orc_graph_neighbors(self.h, v, out.ctypes.data, d)
        return out[:d].copy()

    def closeness_argmax(self):
        out = np.zeros(self.nv, dtype=np.int64)
        n = lib().orc_graph_closeness_argmax(self.h, out.ctypes.data, self.nv)
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_graph_free(self.h)
            self.h = None


class Run:
    """One replicate of the oracle."""

    def __init__(self, graph: Graph, params, seed: int, stream: int, *, candidates=None, faithful=True,
                 want=("events", "clones", "lattice", "lists", "snapshots", "series")):
        p, keep = params if isinstance(params, tuple) else (params, None)
        cand = None
        ncand = 0
        if candidates is not None:
            cand_arr = np.ascontiguousarray(candidates, dtype=np.int64)
            cand, ncand = cand_arr.ctypes.data, len(cand_arr)
        L = lib()
        h = L.orc_run(graph.h, C.byref(p), seed, stream, cand, ncand, 1 if faithful else 0)
        try:
            self.summary = np.frombuffer(C.string_at(L.orc_summary(h), 64), dtype=SUMMARY_DTYPE)[0].copy()
            ptr = C.c_void_p()

            def grab(fn, dtype):
                n = fn(h, C.byref(ptr))
                if n == 0:
                    return np.zeros(0, dtype=dtype)
                return np.frombuffer(C.string_at(ptr.value, n * dtype.itemsize), dtype=dtype).copy()

            if "events" in want:
                self.events = grab(L.orc_events, EVENT_DTYPE)
            if "clones" in want:
                self.clones = grab(L.orc_clones, CLONE_DTYPE)
            if "snapshots" in want:
                self.snapshots = grab(L.orc_snapshots, SNAPSHOT_DTYPE)
            if "series" in want:
                self.series = grab(L.orc_series, np.dtype("<u4"))
            if "lattice" in want:
                self.clone_of_site = np.zeros(graph.nv, dtype=np.uint16)
                self.cell_of_site = np.zeros(graph.nv, dtype=np.uint32)
                L.orc_lattice(h, self.clone_of_site.ctypes.data, self.cell_of_site.ctypes.data)
            if "lists" in want:
                self.lists = []
                for li in range(int(self.summary["n_clones"]) + 1):
                    n = L.orc_list(h, li, C.byref(ptr))
                    a = (np.frombuffer(C.string_at(ptr.value, n * 4), dtype="<i4").copy()
                         if n else np.zeros(0, dtype="<i4"))
                    self.lists.append(a.astype(np.uint32))
        finally:
            L.orc_free(h)
