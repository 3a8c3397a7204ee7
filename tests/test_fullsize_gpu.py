"""BASELINE.json's configurations at FULL size on the GPU, checked through size-independent
properties on MANY replicates (the direct bit-for-bit comparison with the oracle at these sizes is
tests/test_fullsize_parity_gpu.py; the properties here cost nothing per replicate and do not need it):
  * replay round trip: applying the event log to the seeded lattice reproduces the final lattice,
    the clone table counts and n_alive (log -> state is the consumer's contract)
  * summary consistency and genealogy well-formedness
  * results do not depend on how the ensemble is sharded or scheduled (stream g == replicate g)
plus the second correctness leg of the north star: ensemble statistics of the GPU path agree with
the oracle's under a two-sample KS test at alpha = 0.01 (independent streams)."""
import os

import numpy as np
import pytest
from scipy import stats

from parity import TREE_EDGES, TREE_RATES

pytestmark = pytest.mark.gpu

DEFAULT = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4,
               adv_std=0.1, rule="contact")


def replay_lattice(nv, events, seed_sites):
    """Final (clone, cell id) per vertex from the log; seeds = [(vertex, id)]."""
    clone = np.zeros(nv, dtype=np.uint16)
    cid = np.zeros(nv, dtype=np.uint32)
    for v, i in seed_sites:
        clone[v - 1], cid[v - 1] = 1, i
    t, site, site2 = events["type"], events["site"].astype(np.int64) - 1, events["site2"].astype(np.int64) - 1
    cl, ce = events["clone"], events["cell"]
    for k in range(len(events)):
        if t[k] == 2:
            clone[site[k]] = 0
            cid[site[k]] = 0
        elif t[k] == 3:
            clone[site2[k]] = 0
            cid[site2[k]] = 0
            clone[site[k]], cid[site[k]] = cl[k], ce[k]
        else:
            clone[site[k]], cid[site[k]] = cl[k], ce[k]
    return clone, cid


def check_replicates(jsb, graph, point, reps, seed, seeds_of=None, n_check=2):
    from jspace_b200 import _lib as L
    res = jsb.run_ensemble(graph, point, reps, seed=seed, flags=L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS)
    s = res.summaries
    assert np.all((s["status"] == L.ST_OK)), np.bincount(s["status"])
    assert np.all(s["n_alive"] == 0) or s["n_alive"].max() <= graph.nv
    order = np.argsort(-s["n_events"].astype(np.int64))[:n_check]  # the heaviest replicates
    for i in order.tolist():
        ev, cl = res.events(i), res.clones(i)
        clone, cid = res.lattice(i)
        assert len(ev) == s["n_events"][i]
        seeds = seeds_of(i) if seeds_of else [(int(graph.seed_candidates()[0]), 1)]
        rc, ri = replay_lattice(graph.nv, ev, seeds)
        assert np.array_equal(rc, clone) and np.array_equal(ri, cid)
        assert int(np.count_nonzero(clone)) == s["n_alive"][i] == int(cl["count"].sum())
        assert np.array_equal(np.bincount(clone, minlength=len(cl) + 1)[1:], cl["count"])
        alive = res.list(i, 0)
        assert sorted(alive.tolist()) == (np.nonzero(clone)[0] + 1).tolist()
        born = ev[ev["type"] <= 1]
        assert len(np.unique(born["cell"])) == len(born) == 2 * s["n_births"][i]
        assert np.all(born["parent"] < born["cell"])  # a parent id is always older than its children
        assert s["n_effective"][i] == int((ev["flags"] & 1).sum()) or point.t_bottleneck
    out = s.copy()
    res.close()
    return out


def test_c2_helper_warps_leave_full_size_logs_untouched(jsb, monkeypatch):
    """The plain exact loop (no helper warps) against the loop with helpers.  With 24 replicates on 148 SMs every replicate runs with lookahead helpers and,
    once it has 64 clones, on the decoupled clock: the event logs (times bit for bit, patched into the records by
    the clock helper) and the summaries must equal those of the plain loop."""
    from jspace_b200 import _lib as L
    g = jsb.Graph.lattice(200, 200, 2)
    pt = jsb.Point(**DEFAULT, t_bottleneck=[50.0], ratio_bottleneck=[0.8], t_snapshot=[120.0])
    a = jsb.run_ensemble(g, pt, 24, seed=1234, flags=L.OUT_EVENTS)
    monkeypatch.setenv("JSB_NO_HELPERS", "1")
    b = jsb.run_ensemble(g, pt, 24, seed=1234, flags=L.OUT_EVENTS)
    monkeypatch.delenv("JSB_NO_HELPERS")
    sa, sb = a.summaries, b.summaries
    assert sa.tobytes() == sb.tobytes()
    assert sa["n_clones"].max() > 300 and sa["n_iterations"].max() > 2e6
    for i in np.argsort(-sa["n_events"].astype(np.int64))[:4].tolist():
        ea, eb = a.events(i), b.events(i)
        assert ea.tobytes() == eb.tobytes()
        normal = ea[(ea["flags"] & 4) == 0]["time"]  # excised cells are logged at the excision time
        assert np.all(np.diff(normal) >= 0) and normal[-1] <= sa["t_final"][i] <= 200.0 * (1 + 1e-12) + 1.0
    a.close()
    b.close()


def test_c1_100x100_single_replicate_full(jsb):
    g = jsb.Graph.lattice(100, 100, 2)
    s = check_replicates(jsb, g, jsb.Point(**DEFAULT, t_bottleneck=[50.0], ratio_bottleneck=[0.8]), 4, 1234, n_check=4)
    assert s["n_iterations"].max() > 5e5


def test_c2_200x200_ensemble_members_full(jsb):
    g = jsb.Graph.lattice(200, 200, 2)
    s = check_replicates(jsb, g, jsb.Point(**DEFAULT, t_bottleneck=[50.0], ratio_bottleneck=[0.8]), 24, 1234)
    assert s["n_iterations"].max() > 2e6 and s["n_clones"].max() > 300


def test_c3_3d_tree_excision_full(jsb):
    g = jsb.Graph.lattice(350, 350, 3)  # trunc(122500^(1/3)) + 1 = 50 -> 50^3 sites
    assert g.nv == 125000
    for rule in ("contact", "hvoter"):
        p = jsb.Point(method=2, Tf=200.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, rule=rule,
                      tree_edges=TREE_EDGES, node_rate=TREE_RATES, root_node=0, root_rate=0.2,
                      t_bottleneck=[100.0], ratio_bottleneck=[0.8])
        s = check_replicates(jsb, g, p, 16, 1234)
        assert s["n_clones"].max() <= 4 and s["n_excised"].max() > 100


def test_c4_100k_node_geometric_graph_full(jsb):
    from graphs import random_geometric_csr, geometric_centre
    rowptr, colidx, pts = random_geometric_csr(100000, seed=4321, return_points=True)
    g = jsb.Graph.csr(rowptr, colidx)
    assert g.nv > 90000
    # stand-in for the max-closeness vertex of src/J_Space.jl:83-85: the vertex nearest to the
    # middle of the unit square; jsb_graph_seed_candidates computes the exact arg-max when none is supplied
    centre = geometric_centre(pts)
    g.set_seed_candidates([centre])
    s = check_replicates(jsb, g, jsb.Point(**DEFAULT), 8, 1234, seeds_of=lambda i: [(centre, 1)])
    assert s["n_iterations"].max() > 1e5


def test_c5_parameter_sweep_sharded(jsb):
    # BASELINE config 5 at reduced replicate count: the birth/death/advantage grid,
    # run once whole and once as 4 shards; summaries and histograms must be identical
    from jspace_b200 import _lib as L
    g = jsb.Graph.lattice(100, 100, 2)
    pts = [jsb.Point(Tf=200.0, rate_birth=b, rate_death=d, rate_migration=0.01, mu_driver=1e-5, adv_mean=a, adv_std=0.1)
           for b in np.linspace(0.1, 0.6, 8) for d in (0.005, 0.01, 0.02, 0.04) for a in (0.1, 0.2, 0.4, 0.8)]
    reps = 8
    whole = jsb.run_ensemble(g, pts, reps, seed=99)
    total = len(pts) * reps
    parts = []
    hist = np.zeros(64, dtype=np.uint64)
    for k in range(4):
        a, b = jsb.shard_range(k, 4, total)
        r = jsb.run_ensemble(g, pts, reps, seed=99, g_begin=a, g_end=b, warps_per_sm=4 + 2 * k)
        parts.append(r.summaries.copy())
        for p in range(len(pts)):
            if a < (p + 1) * reps and b > p * reps:
                hist += r.histogram(p, 0) if p == 5 else 0
        r.close()
    assert np.concatenate(parts).tobytes() == whole.summaries.tobytes()
    assert np.array_equal(hist, whole.histogram(5, 0))
    assert np.all(whole.summaries["status"] == L.ST_OK)
    # faster birth -> more iterations, on average
    it = whole.summaries["n_iterations"].reshape(8, -1).mean(1)
    assert it[-1] > it[0]
    whole.close()


def test_ensemble_statistics_ks(jsb, oracle):
    """GPU streams 0..N-1 vs oracle streams 10000..10000+N-1: same law, independent randomness."""
    n = 300
    params = dict(Tf=80.0, rate_birth=0.25, rate_death=0.02, rate_migration=0.01, mu_driver=0.03, adv_mean=0.4,
                  adv_std=0.1, rule="contact")
    g = jsb.Graph.lattice(40, 40, 2)
    res = jsb.run_ensemble(g, jsb.Point(**params), n, seed=4242)
    go = oracle.Graph.lattice(2, 40, 40)
    osum = np.array([oracle.Run(go, oracle.make_params(**params), 4242, 10000 + i, faithful=False,
                                want=("clones",)).summary for i in range(n)])
    gs = res.summaries
    for field in ("n_alive", "n_clones", "n_iterations", "n_births"):
        p = stats.ks_2samp(gs[field].astype(float), osum[field].astype(float)).pvalue
        assert p > 0.01, (field, p)
    # largest-clone fraction among surviving replicates
    frac_gpu = [res.clones(i)["count"].max() / gs["n_alive"][i] for i in range(n) if gs["n_alive"][i] > 0]
    res.close()
    frac_cpu = []
    for i in range(n):
        r = oracle.Run(go, oracle.make_params(**params), 4242, 10000 + i, faithful=False, want=("clones",))
        if r.summary["n_alive"] > 0:
            frac_cpu.append(r.clones["count"].max() / r.summary["n_alive"])
    assert stats.ks_2samp(frac_gpu, frac_cpu).pvalue > 0.01


def test_streamed_outputs_equal_outputs_gathered_after_the_kernel(jsb, monkeypatch):
    # From the second call on, the logs of finished replicates are packed and copied to the host while the kernel is
    # still running (jsb_api.cu, "streamed outputs"): rows land in finish order in the host block, the accessors must
    # return exactly what the gather-after-the-kernel path returns, for every replicate.
    from jspace_b200 import _lib as L
    g = jsb.Graph.lattice(100, 100, 2)
    p = jsb.Point(**dict(DEFAULT, Tf=120.0, t_bottleneck=[50.0], ratio_bottleneck=[0.8], t_snapshot=[30.0]))
    flags = L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS

    def run(seed):
        return jsb.run_ensemble(g, p, 1600, seed=seed, flags=flags, log_pool_bytes=2 << 30)
    run(7).close()                      # leaves a pinned block and a gather buffer behind
    a = run(8)                          # streamed
    monkeypatch.setenv("JSB_NO_STREAM_OUT", "1")
    b = run(8)                          # gathered and copied after the kernel
    assert a.launches < b.launches, "the second call was expected to stream its outputs (no gather launch)"
    assert a.summaries.tobytes() == b.summaries.tobytes()
    assert int(a.summaries["n_events"].sum()) > 1_000_000
    for i in range(a.count):
        assert a.events(i).tobytes() == b.events(i).tobytes(), i
        la, lb = a.lattice(i), b.lattice(i)
        assert la[0].tobytes() == lb[0].tobytes() and la[1].tobytes() == lb[1].tobytes(), i
    for i in (0, 5, 1599):
        assert a.list(i, 0).tobytes() == b.list(i, 0).tobytes() and a.clones(i).tobytes() == b.clones(i).tobytes()
        assert a.snapshots(i).tobytes() == b.snapshots(i).tobytes()
    a.close()
    b.close()
