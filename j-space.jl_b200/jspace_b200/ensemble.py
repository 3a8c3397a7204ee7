"""Ensemble interface over the C ABI: graphs, sweep points, jsb_run, results.

`Graph` replaces the MetaGraph returned by the reference's `spatial_graph`
(src/J_Space.jl:43-71); `run_ensemble` runs any number of sweep points × replicates of
`simulate_evolution` (src/J_Space.jl:638-842, :848-1104) on one GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _lib as L

RULE_BY_NAME = {"contact": L.RULE_CONTACT, "voter": L.RULE_VOTER, "h_voter": L.RULE_HVOTER,
                "hvoter": L.RULE_HVOTER}


def rule_code(model) -> int:
    """`model::String` of simulate_evolution; any unknown string never rejects (:767-786)."""
    if isinstance(model, str):
        return RULE_BY_NAME.get(model, L.RULE_NONE)
    return int(model)


class Graph:
    """Handle to a jsb_graph (lattice or CSR)."""

    def __init__(self, handle, kind, meta=None):
        self._h = handle
        self.kind = kind
        self.meta = meta or {}
        self.nv = L.lib().jsb_graph_nv(handle)
        self.ne = L.lib().jsb_graph_ne(handle)

    # -- constructors ------------------------------------------------------------------------
    @classmethod
    def lattice(cls, row: int, col: int, dim: int = 2) -> "Graph":
        h = C.c_void_p()
        L.check(L.lib().jsb_graph_lattice(int(dim), int(row), int(col), C.byref(h)))
        return cls(h, "lattice", {"row": row, "col": col, "dim": dim})

    @classmethod
    def csr(cls, rowptr, colidx) -> "Graph":
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(colidx, dtype=np.int32)
        h = C.c_void_p()
        L.check(L.lib().jsb_graph_csr(len(rp) - 1, rp.ctypes.data, ci.ctypes.data, C.byref(h)))
        return cls(h, "csr")

    @classmethod
    def dense_file(cls, path) -> "Graph":
        h = C.c_void_p()
        L.check(L.lib().jsb_graph_dense_file(os.fsencode(path), C.byref(h)))
        return cls(h, "csr", {"path": str(path)})

    # -- queries -----------------------------------------------------------------------------
    def neighbors(self, site: int) -> np.ndarray:
        out = np.zeros(16, dtype=np.int64)
        d = L.lib().jsb_graph_neighbors(self._h, int(site), out.ctypes.data, len(out))
        if d < 0:
            L.check(d)
        if d > len(out):
            out = np.zeros(d, dtype=np.int64)
            L.lib().jsb_graph_neighbors(self._h, int(site), out.ctypes.data, d)
        return out[:d].copy()

    def seed_candidates(self) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_int64()
        L.check(L.lib().jsb_graph_seed_candidates(self._h, C.byref(p), C.byref(n)))
        return L.view(p.value, n.value, np.int64)

    def set_seed_candidates(self, sites: Sequence[int]) -> None:
        a = np.ascontiguousarray(sites, dtype=np.int64)
        L.check(L.lib().jsb_graph_set_seed_candidates(self._h, a.ctypes.data, len(a)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                L.lib().jsb_graph_free(h)
            except Exception:
                pass
            self._h = None


@dataclass
class Point:
    """One sweep point = the arguments of one simulate_evolution call."""
    method: int = L.METHOD_RANDOM
    Tf: float = 200.0
    rate_birth: float = 0.2
    rate_death: float = 0.01
    rate_migration: float = 0.01
    mu_driver: float = 0.01
    adv_mean: float = 0.4
    adv_std: float = 0.1
    rule: object = "contact"
    n_start_cells: int = 1
    t_bottleneck: Sequence[float] = ()
    ratio_bottleneck: Sequence[float] = ()
    t_snapshot: Sequence[float] = ()
    tree_edges: Sequence[tuple] = ()   # (parent node, child node), file order, 0-based
    node_rate: Sequence[float] = ()
    root_node: int = 0
    root_rate: float = 0.0
    quirk_flags: int = 0
    max_iterations: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def to_struct(self) -> L.Params:
        if len(self.t_bottleneck) != len(self.ratio_bottleneck):
            raise ValueError("t_bottleneck and ratio_bottleneck must have the same length")
        p = L.Params()
        p.size = C.sizeof(L.Params)
        p.method = self.method
        p.Tf, p.rate_birth, p.rate_death = float(self.Tf), float(self.rate_birth), float(self.rate_death)
        p.rate_migration, p.mu_driver = float(self.rate_migration), float(self.mu_driver)
        p.adv_mean, p.adv_std = float(self.adv_mean), float(self.adv_std)
        p.rule = rule_code(self.rule)
        p.n_start_cells = int(self.n_start_cells)
        keep = self._keep
        keep.clear()

        def darr(v):
            a = np.ascontiguousarray(np.asarray(list(v), dtype=np.float64))
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double)) if a.size else None

        def iarr(v):
            a = np.ascontiguousarray(np.asarray(list(v), dtype=np.int32))
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32)) if a.size else None

        p.n_bottleneck = len(self.t_bottleneck)
        p.t_bottleneck = darr(self.t_bottleneck)
        p.ratio_bottleneck = darr(self.ratio_bottleneck)
        p.n_snapshots = len(self.t_snapshot)
        p.t_snapshot = darr(self.t_snapshot)
        p.n_tree_nodes = len(self.node_rate)
        p.n_tree_edges = len(self.tree_edges)
        p.edge_parent = iarr([e[0] for e in self.tree_edges])
        p.edge_child = iarr([e[1] for e in self.tree_edges])
        p.node_rate = darr(self.node_rate)
        p.root_node = int(self.root_node)
        p.root_rate = float(self.root_rate)
        p.quirk_flags = int(self.quirk_flags)
        p.max_iterations = int(self.max_iterations)
        return p


def series_from_rows(rows, rows_before_snapshot=(), n0: int = 1):
    """jsb_series_row records -> (list_len_node_occ, CA_Alive_TOT).  One push per run of rows that ends with a row
    without SERIES_MORE; `gain` beyond the number of clones seen so far is a new clone of one cell; a snapshot
    inserts a copy of the current counts before the push at its row position (src/J_Space.jl:712)."""
    n_occ = [int(n0)]
    counts = [int(n0)] if n0 else []
    ca_tot = []
    snap_at = sorted(int(x) for x in rows_before_snapshot)
    si = 0
    k = 0
    while k < len(rows):
        while si < len(snap_at) and snap_at[si] <= k:
            ca_tot.append(list(counts))
            si += 1
        while True:
            r = rows[k]
            if r["loss"]:
                counts[int(r["loss"]) - 1] -= 1
            if r["gain"]:
                g = int(r["gain"])
                if g > len(counts):
                    counts.append(1)
                else:
                    counts[g - 1] += 1
            k += 1
            if not (int(r["n_alive"]) & L.SERIES_MORE):
                break
        n_occ.append(int(r["n_alive"]) & 0x7FFFFFFF)
        ca_tot.append(list(counts))
    while si < len(snap_at):
        ca_tot.append(list(counts))
        si += 1
    return n_occ, ca_tot


class EnsembleResult:
    """Owns a jsb_result."""

    def __init__(self, handle, graph: Graph, flags: int):
        self._h = handle
        self.graph = graph
        self.flags = flags
        lib = L.lib()
        self.count = lib.jsb_result_count(handle)
        p = C.c_void_p()
        L.check(lib.jsb_result_summaries(handle, C.byref(p)))
        self.summaries = L.view(p.value, self.count, L.SUMMARY_DTYPE)
        ms = C.c_double()
        L.check(lib.jsb_result_kernel_ms(handle, C.byref(ms)))
        self.kernel_ms = ms.value
        nl = C.c_int64()
        L.check(lib.jsb_result_launches(handle, C.byref(nl)))
        self.launches = nl.value

    def clones(self, i: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int32()
        L.check(L.lib().jsb_result_clones(self._h, i, C.byref(p), C.byref(n)))
        return L.view(p.value, n.value, L.CLONE_DTYPE)

    def events(self, i: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int64()
        L.check(L.lib().jsb_result_events(self._h, i, C.byref(p), C.byref(n)))
        return L.view(p.value, n.value, L.EVENT_DTYPE)

    def lattice(self, i: int):
        pc, pi = C.c_void_p(), C.c_void_p()
        L.check(L.lib().jsb_result_lattice(self._h, i, C.byref(pc), C.byref(pi)))
        return L.view(pc.value, self.graph.nv, np.uint16), L.view(pi.value, self.graph.nv, np.uint32)

    def list(self, i: int, which: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int64()
        L.check(L.lib().jsb_result_list(self._h, i, which, C.byref(p), C.byref(n)))
        return L.view(p.value, n.value, np.uint32)

    def snapshots(self, i: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int32()
        L.check(L.lib().jsb_result_snapshots(self._h, i, C.byref(p), C.byref(n)))
        return L.view(p.value, n.value, L.SNAPSHOT_DTYPE)

    def series_rows(self, i: int):
        """JSB_OUT_SERIES: (rows, rows_before_snapshot) of replicate i — one row per push to list_len_node_occ /
        CA_Alive_TOT (an excision removing k cells: k rows, all but the last with SERIES_MORE)."""
        p, n, ps, ns = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int32()
        L.check(L.lib().jsb_result_series(self._h, i, C.byref(p), C.byref(n), C.byref(ps), C.byref(ns)))
        snaps = L.view(ps.value, ns.value, np.int64) if ns.value else np.zeros(0, dtype=np.int64)
        return L.view(p.value, n.value, L.SERIES_DTYPE), snaps

    def series(self, i: int, n0: int = 1):
        """(list_len_node_occ, CA_Alive_TOT) of replicate i as the reference builds them (src/J_Space.jl:656-658,
        712, 740-741, 757-758, 807-808, 835-836), from the device-side series; n0 = number of starting cells."""
        rows, snaps = self.series_rows(i)
        return series_from_rows(rows, snaps, n0)

    def histogram(self, point: int, kind: int) -> np.ndarray:
        out = np.zeros(64, dtype=np.uint64)
        L.check(L.lib().jsb_result_histogram(self._h, point, kind, out.ctypes.data))
        return out

    def genealogy(self, i: int, sample_cells) -> np.ndarray:
        s = np.ascontiguousarray(sample_cells, dtype=np.uint32)
        p, n = C.c_void_p(), C.c_int64()
        L.check(L.lib().jsb_result_genealogy(self._h, i, s.ctypes.data, len(s), C.byref(p), C.byref(n)))
        try:
            return L.view(p.value, n.value, L.RELATION_DTYPE)
        finally:
            L.lib().jsb_free(p)

    def close(self):
        if self._h:
            L.lib().jsb_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_ensemble(graph: Graph, points, reps_per_point: int = 1, *, seed: int = 1234, flags: int = 0,
                 g_begin: int = 0, g_end: int = 0, device: int = 0, warps_per_sm: int = 0,
                 log_pool_bytes: int = 0, devices=None) -> EnsembleResult:
    """jsb_run: replicate g = point*reps_per_point + r uses Philox stream g.  `devices`: several GPUs of the
    box in one call (the replicate range is cut into contiguous shards, one per device)."""
    if isinstance(points, Point):
        points = [points]
    arr = (L.Params * len(points))(*[p.to_struct() for p in points])
    o = L.RunOpts()
    o.size = C.sizeof(L.RunOpts)
    o.flags = flags
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    o.g_begin, o.g_end = int(g_begin), int(g_end)
    o.device = int(device)
    o.warps_per_sm = int(warps_per_sm)
    o.log_pool_bytes = int(log_pool_bytes)
    dev_arr = None
    if devices:
        dev_arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        o.n_devices = len(devices)
        o.devices = C.cast(dev_arr, C.POINTER(C.c_int32))
    h = C.c_void_p()
    L.check(L.lib().jsb_run(graph._h, arr, len(points), int(reps_per_point), C.byref(o), C.byref(h)))
    return EnsembleResult(h, graph, flags)
