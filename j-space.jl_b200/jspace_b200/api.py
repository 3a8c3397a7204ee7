"""Python mirror of the reference's Julia API for the spatial clonal-dynamics path.

Same names, argument order and 7-tuple return as the reference (src/J_Space.jl:25):

    spatial_graph(row, col, seed; dim=2, n_cell=1)           src/J_Space.jl:43
    spatial_graph(path, seed; n_cell=1)                      src/J_Space.jl:67
    simulate_evolution(G, Tf, rate_birth, rate_death, rate_migration, mu_dri,
                       driv_average_advantage, driv_std_advantage, model, seed; ...)   :638
    simulate_evolution(G, Tf, rate_death, rate_migration, mu_dri, model,
                       edge_list_path, driv_adv_path, seed; ...)                        :848
    -> (df, G, list_len_node_occ, set_mut_pop, Gs_plot, CA_Alive_TOT, α)               :841/:1103

The dynamics run on the GPU through libjspace_b200.so; this module only converts arguments and
re-materialises the reference's output objects from the library's records.

Differences that cannot be hidden (also listed in DESIGN.md §6):
  * `seed` is a `PhiloxSeed` (or an int): randomness comes from the counter-based draw contract,
    not from Julia's MersenneTwister.  Every simulate_evolution call on the same PhiloxSeed uses
    the next stream, as successive calls on one MersenneTwister see fresh numbers.
  * Cell ids are sequential integers in `uuid1` call order, not time-based UUIDs.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from . import _lib as L
from .ensemble import Graph, Point, run_ensemble

EVENT_NAMES = {L.EV_DUPLICATE: "Duplicate", L.EV_MUTATION: "Mutation", L.EV_DEATH: "Death",
               L.EV_MIGRATION: "Migration"}


class PhiloxSeed:
    """Stands in for `MersenneTwister(seed)`: a Philox key plus a stream counter."""

    def __init__(self, seed: int, stream: int = 0):
        self.seed = int(seed)
        self.stream = int(stream)

    def next_stream(self) -> int:
        s = self.stream
        self.stream += 1
        return s


def _as_seed(seed) -> PhiloxSeed:
    return seed if isinstance(seed, PhiloxSeed) else PhiloxSeed(int(seed))


class SpatialGraph:
    """What the reference keeps in a MetaGraph: the graph plus per-vertex cell properties."""

    def __init__(self, graph: Graph, n_cell: int = 1):
        self.graph = graph
        self.n_cell = int(n_cell)
        self.nv = graph.nv
        self.clone = None      # uint16 per vertex (0 = empty), :Subpop
        self.cell_id = None    # uint32 per vertex, :id
        self.genotypes = None  # set_mut_pop
        self.alpha = None

    # -- MetaGraphs-like accessors -------------------------------------------------------------
    def cells_alive(self) -> List[int]:
        """cells_alive(G), src/J_Space.jl:296-300 (ascending vertex numbers)."""
        if self.clone is None:
            return []
        return (np.nonzero(self.clone)[0] + 1).tolist()

    def has_prop(self, v: int, key: str) -> bool:
        if key == "name":
            return True
        return self.clone is not None and self.clone[v - 1] != 0 and key in ("id", "mutation", "Subpop", "Fit")

    def get_prop(self, v: int, key: str):
        if key == "name":
            return f"node{v}"
        c = int(self.clone[v - 1])
        if c == 0:
            raise KeyError(key)
        if key == "id":
            return int(self.cell_id[v - 1])
        if key == "Subpop":
            return c
        if key == "Fit":
            return float(self.alpha[c - 1])
        if key == "mutation":
            return self.genotypes[c - 1]
        raise KeyError(key)

    def props(self, v: int) -> dict:
        d = {"name": f"node{v}"}
        if self.clone is not None and self.clone[v - 1] != 0:
            for k in ("id", "mutation", "Subpop", "Fit"):
                d[k] = self.get_prop(v, k)
        return d

    def copy(self) -> "SpatialGraph":
        g = SpatialGraph(self.graph, self.n_cell)
        for k in ("clone", "cell_id", "genotypes", "alpha"):
            v = getattr(self, k)
            setattr(g, k, None if v is None else (v.copy() if hasattr(v, "copy") else list(v)))
        return g


def spatial_graph(*args, dim: int = 2, n_cell: int = 1) -> SpatialGraph:
    """spatial_graph(row, col, seed; dim, n_cell) or spatial_graph(path, seed; n_cell)."""
    if len(args) == 3:
        row, col, _seed = args
        return SpatialGraph(Graph.lattice(int(row), int(col), int(dim)), n_cell)
    if len(args) == 2:
        path, _seed = args
        return SpatialGraph(Graph.dense_file(path), n_cell)
    raise TypeError("spatial_graph(row, col, seed; dim, n_cell) or spatial_graph(path, seed; n_cell)")


# ----------------------------------------------------------------------------------------------
# driver-tree files (method 2), src/J_Space.jl:871-881
# ----------------------------------------------------------------------------------------------
def read_driver_tree(edge_list_path, driv_adv_path):
    """readdlm of both files -> (names, edges as node indices in file order, rates by node,
    root node, rate of the FIRST row).  Mirrors :871-873,:881,:503-504."""
    with open(edge_list_path) as f:
        edges_txt = [ln.split() for ln in f if ln.strip()]
    with open(driv_adv_path) as f:
        adv_txt = [ln.split() for ln in f if ln.strip()]
    if any(len(e) != 2 for e in edges_txt) or any(len(a) != 2 for a in adv_txt):
        raise ValueError("driver tree files must have two whitespace-separated columns")
    names = []
    for a in adv_txt:
        if a[0] not in names:
            names.append(a[0])
    for e in edges_txt:
        for nm in e:
            if nm not in names:
                raise ValueError(f"driver '{nm}' has no birth rate in {driv_adv_path} (reference: BoundsError at :503)")
    idx = {n: i for i, n in enumerate(names)}
    edges = [(idx[p], idx[c]) for p, c in edges_txt]
    rate = [0.0] * len(names)
    seen = set()
    for nm, r in adv_txt:  # findall(...)[1]: the first row with that name wins
        if nm not in seen:
            rate[idx[nm]] = float(r)
            seen.add(nm)
    parents = [p for p, _ in edges_txt]
    children = {c for _, c in edges_txt}
    roots = [p for p in parents if p not in children]  # setdiff!(parents, children)[1]
    if not roots:
        raise ValueError("driver tree has no root")
    return names, edges, rate, idx[roots[0]], float(adv_txt[0][1])


# ----------------------------------------------------------------------------------------------
# re-materialisation of the reference's outputs
# ----------------------------------------------------------------------------------------------
def genotypes_from_clones(clones: np.ndarray, names: Sequence[str] | None = None) -> list:
    """set_mut_pop: method 1 -> 1, [1, 143], ...; method 2 -> "driver_1", ["driver_1", "driver_2"]."""
    out = []
    for i, c in enumerate(clones):
        lab = int(c["label"])
        item = names[lab - 1] if names is not None else lab
        if int(c["parent"]) == 0:
            out.append(item)
        else:
            par = out[int(c["parent"]) - 1]
            out.append((list(par) if isinstance(par, list) else [par]) + [item])
    return out


def replay_series(events: np.ndarray, n0: int, snap_positions=()):
    """list_len_node_occ (initial n :658, then one push per effective event) and CA_Alive_TOT
    (length.(ca_subpop) per effective event plus one extra row per snapshot :712), replayed from
    the log.  Rows flagged JSB_EVF_GROUP_START open one effective event:
      Migration                      -> nothing changes
      Death                          -> its clone and n lose one cell
      [displaced Death] + birth pair -> (Duplicate, Duplicate): the clone gains one cell;
                                        (Mutation, Duplicate): a new clone with one cell appears
      excision Deaths                -> each removes one cell; one push for the whole block
    An excision that removed nothing (k = 0) writes no rows and is therefore not visible here.
    """
    n = int(n0)
    counts = [n] if n else []
    series = [n]
    ca_tot = []
    if len(events) == 0:
        ca_tot.extend([list(counts) for _ in snap_positions])
        return series, ca_tot
    starts = np.nonzero(events["flags"] & L.EVF_GROUP_START)[0]
    ends = np.append(starts[1:], len(events))
    snaps = sorted(int(p) for p in snap_positions)
    si = 0
    types, clones = events["type"], events["clone"]
    for gs, ge in zip(starts.tolist(), ends.tolist()):
        while si < len(snaps) and snaps[si] <= gs:
            ca_tot.append(list(counts))
            si += 1
        for i in range(gs, ge):
            if types[i] == L.EV_DEATH:
                counts[int(clones[i]) - 1] -= 1
                n -= 1
        if ge - gs >= 2 and types[ge - 1] == L.EV_DUPLICATE and types[ge - 2] in (L.EV_DUPLICATE, L.EV_MUTATION):
            if types[ge - 2] == L.EV_MUTATION:
                counts.append(1)
            else:
                counts[int(clones[ge - 2]) - 1] += 1
            n += 1
        series.append(n)
        ca_tot.append(list(counts))
    while si < len(snaps):
        ca_tot.append(list(counts))
        si += 1
    return series, ca_tot


class EventFrame:
    """The reference's `df` (Event, Time, Cell, Notes), src/J_Space.jl:284-290, kept as the
    library's record array; `to_pandas()` builds the DataFrame the reference returns."""

    def __init__(self, records: np.ndarray, genotypes: list, method: int):
        self.records = records
        self.genotypes = genotypes
        self.method = method

    def __len__(self):
        return len(self.records)

    @property
    def Event(self):
        return [EVENT_NAMES[int(t)] for t in self.records["type"]]

    @property
    def Time(self):
        return self.records["time"]

    @property
    def Cell(self):
        return self.records["cell"]

    def notes(self, i: int):
        r = self.records[i]
        t = int(r["type"])
        if t == L.EV_DUPLICATE:
            return [int(r["parent"])]
        if t == L.EV_MUTATION:
            if self.method == L.METHOD_TREE:
                return [int(r["parent"]), self.genotypes[int(r["clone"]) - 1]]  # :508
            return [int(r["parent"]), int(r["label"])]                           # :387
        return None  # `undef`

    def to_pandas(self):
        import pandas as pd
        return pd.DataFrame({"Event": self.Event, "Time": self.Time.copy(), "Cell": self.Cell.copy(),
                             "Notes": [self.notes(i) for i in range(len(self))]})


def _state_from_log(base: SpatialGraph, events: np.ndarray, upto: int, seeds) -> SpatialGraph:
    """Lattice state after the first `upto` log rows (snapshots, :708-714)."""
    g = SpatialGraph(base.graph, base.n_cell)
    clone = np.zeros(base.nv, dtype=np.uint16)
    cid = np.zeros(base.nv, dtype=np.uint32)
    for site, ident in seeds:
        clone[site - 1] = 1
        cid[site - 1] = ident
    for r in events[:upto]:
        t = int(r["type"])
        s = int(r["site"]) - 1
        if t == L.EV_DEATH:
            clone[s] = 0
            cid[s] = 0
        elif t == L.EV_MIGRATION:
            s2 = int(r["site2"]) - 1
            clone[s2] = 0
            cid[s2] = 0
            clone[s] = r["clone"]
            cid[s] = r["cell"]
        else:
            clone[s] = r["clone"]
            cid[s] = r["cell"]
    g.clone, g.cell_id = clone, cid
    return g


def simulate_evolution(G: SpatialGraph, Tf, *args, Time_of_sampling=(), t_bottleneck=(), ratio_bottleneck=(),
                       device: int = 0, quirk_flags: int = 0):
    """Both reference methods; dispatch on the number of positional arguments like Julia's
    multiple dispatch: 10 positional -> random drivers (:638), 9 -> driver tree (:848)."""
    names = None
    if len(args) == 8:
        rate_birth, rate_death, rate_migration, mu_dri, adv_avg, adv_std, model, seed = args
        point = Point(method=L.METHOD_RANDOM, Tf=float(Tf), rate_birth=rate_birth, rate_death=rate_death,
                      rate_migration=rate_migration, mu_driver=mu_dri, adv_mean=adv_avg, adv_std=adv_std, rule=model)
    elif len(args) == 7:
        rate_death, rate_migration, mu_dri, model, edge_list_path, driv_adv_path, seed = args
        names, edges, rate, root, root_rate = read_driver_tree(edge_list_path, driv_adv_path)
        point = Point(method=L.METHOD_TREE, Tf=float(Tf), rate_death=rate_death, rate_migration=rate_migration,
                      mu_driver=mu_dri, rule=model, tree_edges=edges, node_rate=rate, root_node=root,
                      root_rate=root_rate)
    else:
        raise TypeError("simulate_evolution: expected 10 (random drivers) or 9 (driver tree) positional arguments")
    point.n_start_cells = G.n_cell
    point.t_bottleneck = list(t_bottleneck)
    point.ratio_bottleneck = list(ratio_bottleneck)
    point.t_snapshot = [float(x) for x in Time_of_sampling]
    point.quirk_flags = quirk_flags
    ps = _as_seed(seed)
    stream = ps.next_stream()
    res = run_ensemble(G.graph, point, stream + 1, seed=ps.seed, flags=L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS,
                       g_begin=stream, g_end=stream + 1, device=device)
    try:
        summ = res.summaries[0]
        if int(summ["status"]) == L.ST_NONTERMINATING:
            import warnings
            warnings.warn("rate_death == 0 and the lattice filled: the reference loop (src/J_Space.jl:687-688) "
                          "never terminates; stopped at Tf")
        elif int(summ["status"]) not in (L.ST_OK, L.ST_MAX_ITER):
            raise L.JSpaceError(-4, f"replicate stopped with status {int(summ['status'])} "
                                    "(2: all 999 driver labels used — the reference throws at src/J_Space.jl:371)")
        events = res.events(0)
        clones = res.clones(0)
        clone_of_site, cell_of_site = res.lattice(0)
        snaps = res.snapshots(0)
    finally:
        res.close()
    genotypes = genotypes_from_clones(clones, names)
    alpha = [float(a) for a in clones["alpha"]]
    out = G.copy()
    out.clone, out.cell_id, out.genotypes, out.alpha = clone_of_site, cell_of_site, genotypes, alpha
    n0 = G.n_cell if G.n_cell != 1 else (1 if len(G.graph.seed_candidates()) > 0 else 0)
    series, ca_tot = replay_series(events, n0, [int(s["n_events_before"]) for s in snaps])
    df = EventFrame(events, genotypes, point.method)
    gs_plot = []
    if len(snaps):
        init_clone, init_id = _initial_lattice(G, point, ps.seed, stream, device)
        seeds = [(int(v) + 1, int(init_id[v])) for v in np.nonzero(init_clone)[0]]
        for s in snaps:
            g = _state_from_log(G, events, int(s["n_events_before"]), seeds)
            g.genotypes, g.alpha = genotypes, alpha
            gs_plot.append(g)
    return df, out, series, genotypes, gs_plot, ca_tot, alpha


def _initial_lattice(G: SpatialGraph, point: Point, seed: int, stream: int, device: int):
    """The seeded lattice (initialize_graph, src/J_Space.jl:79-163): the same replicate with
    Tf = 0, whose loop is never entered."""
    import copy
    p0 = copy.copy(point)
    p0._keep = []
    p0.Tf = 0.0
    res = run_ensemble(G.graph, p0, stream + 1, seed=seed, flags=L.OUT_LATTICE, g_begin=stream, g_end=stream + 1,
                       device=device)
    try:
        return res.lattice(0)
    finally:
        res.close()
