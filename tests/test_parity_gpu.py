"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle, bit-exact, on seeded inputs at
sizes the oracle finishes in seconds.  Integer fields (cell ids, parents, sites, clones, labels,
event types) must be identical; fp64 times and fitness values are required to be identical too
(the north-star tolerance is 1e-12 relative; we hold the stricter bar because both sides
implement the same IEEE expression order)."""
import os

import numpy as np
import pytest

from parity import TREE_EDGES, TREE_RATES, compare_runs

pytestmark = pytest.mark.gpu

DEFAULT = dict(Tf=60.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01,
               adv_mean=0.4, adv_std=0.1, rule="contact")
TREE = dict(method=2, Tf=80.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.05, rule="contact",
            tree_edges=TREE_EDGES, node_rate=TREE_RATES, root_node=0, root_rate=0.2)


def test_contact_2d_default(jsb, oracle):
    s = compare_runs(jsb, oracle, ("lattice", 2, 40, 40), DEFAULT, reps=8)
    assert sum(int(x["n_births"]) for x in s) > 1000


def test_c1_100x100_full_run(jsb, oracle):
    # BASELINE config 1: 100x100, Parameters.toml dynamics, one excision
    p = dict(DEFAULT, Tf=200.0, t_bottleneck=[50.0], ratio_bottleneck=[0.8])
    s = compare_runs(jsb, oracle, ("lattice", 2, 100, 100), p, reps=2)
    assert max(int(x["n_clones"]) for x in s) > 30


@pytest.mark.parametrize("rule", ["voter", "hvoter", "anything-else"])
def test_rules(jsb, oracle, rule):
    p = dict(DEFAULT, rule=rule, Tf=40.0 if rule != "contact" else 60.0, mu_driver=0.05)
    compare_runs(jsb, oracle, ("lattice", 2, 24, 24), p, reps=6)


def test_serial_debug_path_matches(jsb, oracle):
    from jspace_b200 import _lib as L
    p = dict(DEFAULT, Tf=45.0, quirk_flags=L.QUIRK_DEBUG_SERIAL)
    compare_runs(jsb, oracle, ("lattice", 2, 30, 30), p, reps=4)


@pytest.mark.parametrize("flag", ["QUIRK_DEBUG_EXACT_CLASSES", "QUIRK_DEBUG_WIDE_GUARD"])
def test_class_selection_fallback_paths(jsb, oracle, flag):
    # the exact prob_cum table (reference cumsum order) must give the same trajectory as the
    # guarded k*lambda comparison, including with > 127 clones (pairwise cumsum splits)
    from jspace_b200 import _lib as L
    p = dict(DEFAULT, Tf=45.0, quirk_flags=getattr(L, flag))
    compare_runs(jsb, oracle, ("lattice", 2, 30, 30), p, reps=4)
    p = dict(DEFAULT, Tf=110.0, mu_driver=0.2, adv_mean=0.02, adv_std=0.01, quirk_flags=getattr(L, flag))
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, reps=1)


def test_decoupled_clock_helper_paths(jsb, oracle):
    # JSB_QUIRK_DEBUG_DECOUPLE: the exact fold and time chain run on a helper warp one round behind
    # the replicate, which works on an approximate table/clock (by default only with >= 64 clones);
    # event times, fitness values and every decision must not change.  Covers displacement (voter
    # rules), driver births, the driver tree, excisions and snapshots (exact-clock hand-backs),
    # extinction, and a run long enough for the periodic re-base of the approximate table.
    from jspace_b200 import _lib as L
    q = L.QUIRK_DEBUG_DECOUPLE
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), dict(DEFAULT, quirk_flags=q), reps=8)
    for rule in ("voter", "hvoter", "anything-else"):
        compare_runs(jsb, oracle, ("lattice", 2, 24, 24), dict(DEFAULT, rule=rule, Tf=40.0, mu_driver=0.05, quirk_flags=q), reps=6)
    compare_runs(jsb, oracle, ("lattice", 3, 30, 30), dict(TREE, rule="hvoter", t_bottleneck=[40.0], ratio_bottleneck=[0.8],
                                                          quirk_flags=q), reps=4)
    compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(DEFAULT, Tf=70.0, t_bottleneck=[30.0, 50.0], ratio_bottleneck=[0.5, 0.9],
                                                          t_snapshot=[10.0, 35.0, 69.0], quirk_flags=q | L.QUIRK_EXCISION_INTENT), reps=3)
    compare_runs(jsb, oracle, ("lattice", 2, 30, 30), dict(DEFAULT, Tf=60.0, rate_death=0.15, quirk_flags=q), reps=6)  # extinctions
    compare_runs(jsb, oracle, ("lattice", 2, 12, 12), dict(DEFAULT, max_iterations=5000, quirk_flags=q), reps=4)
    p = dict(DEFAULT, Tf=130.0, mu_driver=0.2, adv_mean=0.02, adv_std=0.01, quirk_flags=q)  # many clones, > 1024 events
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, reps=2)
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), dict(p, quirk_flags=q | L.QUIRK_DEBUG_WIDE_GUARD), reps=1)  # constant hand-backs
    # 32-bit site ids (nv > 65535) and a CSR graph
    compare_runs(jsb, oracle, ("lattice", 2, 300, 300), dict(DEFAULT, Tf=45.0, rate_death=0.03, rate_migration=0.05, quirk_flags=q), reps=2)
    from graphs import random_geometric_csr
    rowptr, colidx = random_geometric_csr(3000, seed=4321)
    compare_runs(jsb, oracle, ("csr", rowptr, colidx), dict(DEFAULT, Tf=80.0, quirk_flags=q), reps=3, candidates=[17])


def test_helper_warps_do_not_change_results(jsb, oracle, monkeypatch):
    # the same replicates with and without lookahead / clock helpers (JSB_NO_HELPERS, JSB_NO_CLOCK)
    p = dict(DEFAULT, Tf=110.0, mu_driver=0.1)
    for env in ("JSB_NO_HELPERS", "JSB_NO_CLOCK"):
        monkeypatch.setenv(env, "1")
        compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, reps=3)
        monkeypatch.delenv(env)


def test_large_lattice_uint32_sites(jsb, oracle):
    # nv > 65536 switches the ordered lists to 32-bit site ids
    p = dict(DEFAULT, Tf=45.0, rate_death=0.03, rate_migration=0.05)
    compare_runs(jsb, oracle, ("lattice", 2, 300, 300), p, reps=3)
    p = dict(DEFAULT, Tf=40.0, rule="voter", mu_driver=0.05)
    compare_runs(jsb, oracle, ("lattice", 3, 270, 270), p, reps=2)


def test_3d_tree_excision(jsb, oracle):
    # BASELINE config 3 at reduced size: 3D cube, driver tree, hvoter and contact, one excision
    for rule in ("contact", "hvoter"):
        p = dict(TREE, rule=rule, t_bottleneck=[40.0], ratio_bottleneck=[0.8])
        compare_runs(jsb, oracle, ("lattice", 3, 30, 30), p, reps=4)


def test_excision_branches_and_quirks(jsb, oracle):
    from jspace_b200 import _lib as L
    base = dict(DEFAULT, Tf=70.0)
    # two excisions: reference method 1 only fires the first; intent mode fires both
    compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(base, t_bottleneck=[30.0, 50.0], ratio_bottleneck=[0.5, 0.9]), reps=3)
    compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(base, t_bottleneck=[30.0, 50.0], ratio_bottleneck=[0.5, 0.9],
                                                          quirk_flags=L.QUIRK_EXCISION_INTENT), reps=3)
    # small ratios reach StatsBase's k == 1, k == 2 and self-avoid branches
    seen = set()
    for ratio in (0.01, 0.02, 0.03, 0.05, 0.0):
        out = compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(base, t_bottleneck=[60.0], ratio_bottleneck=[ratio]), reps=6)
        seen |= {int(x["n_excised"]) for x in out}
    assert {0, 1, 2} <= seen and max(seen) >= 3
    # method 2 advances the excision index every iteration (only the last entry can fire late)
    compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(TREE, t_bottleneck=[20.0, 45.0], ratio_bottleneck=[0.5, 0.7]), reps=3)
    compare_runs(jsb, oracle, ("lattice", 2, 32, 32), dict(TREE, t_bottleneck=[20.0, 45.0], ratio_bottleneck=[0.5, 0.7],
                                                          quirk_flags=L.QUIRK_EXCISION_INTENT), reps=3)


def test_snapshots_and_multi_seed_cells(jsb, oracle):
    p = dict(DEFAULT, Tf=50.0, t_snapshot=[10.0, 20.0, 49.0, 500.0], n_start_cells=5)
    compare_runs(jsb, oracle, ("lattice", 2, 30, 30), p, reps=4)


def test_empty_and_degenerate_inputs(jsb, oracle):
    compare_runs(jsb, oracle, ("lattice", 2, 10, 10), dict(DEFAULT, n_start_cells=0), reps=2)
    compare_runs(jsb, oracle, ("lattice", 2, 1, 1), dict(DEFAULT), reps=2)          # single site
    compare_runs(jsb, oracle, ("lattice", 2, 1, 7), dict(DEFAULT, Tf=100.0), reps=3)  # a path
    compare_runs(jsb, oracle, ("lattice", 2, 6, 6), dict(DEFAULT, Tf=0.0), reps=1)  # loop never entered
    # rate_death == 0 on a lattice that fills: the reference loop never exits -> flagged
    s = compare_runs(jsb, oracle, ("lattice", 2, 5, 5), dict(DEFAULT, rate_death=0.0, Tf=400.0), reps=2)
    assert all(int(x["status"]) == 1 for x in s)
    s = compare_runs(jsb, oracle, ("lattice", 2, 12, 12), dict(DEFAULT, max_iterations=100), reps=4)
    assert any(int(x["status"]) == 4 for x in s)


def test_adjacency_graph(jsb, oracle):
    path = os.path.join(os.path.dirname(__file__), "golden", "Example_adj_matrix.txt")
    compare_runs(jsb, oracle, ("dense", path), dict(DEFAULT, Tf=100.0, rate_death=0.05), reps=8)


def test_random_geometric_graph_csr(jsb, oracle):
    # BASELINE config 4 at reduced size
    from graphs import random_geometric_csr
    rowptr, colidx = random_geometric_csr(3000, seed=4321)
    compare_runs(jsb, oracle, ("csr", rowptr, colidx), dict(DEFAULT, Tf=80.0), reps=4, candidates=[17])


def test_stream_offset_equals_sharding(jsb, oracle):
    # replicate g uses stream g wherever it runs: a shard starting at g=5 reproduces streams 5..8
    compare_runs(jsb, oracle, ("lattice", 2, 24, 24), dict(DEFAULT, Tf=40.0), reps=4, g_begin=5)


def test_many_clones_global_class_table(jsb, oracle):
    # > 126 clones pushes the class table out of shared memory and through the pairwise cumsum
    p = dict(DEFAULT, Tf=130.0, mu_driver=0.2, adv_mean=0.02, adv_std=0.01)
    s = compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, reps=2)
    assert max(int(x["n_clones"]) for x in s) > 257  # at least one pairwise split of the cumsum


def test_arena_compaction(jsb, oracle, monkeypatch):
    # voter turnover with a minimal list arena forces in-place compactions
    monkeypatch.setenv("JSB_DEBUG_ARENA_ENTRIES", "12000")
    p = dict(DEFAULT, rule="voter", Tf=150.0, mu_driver=0.01, rate_death=0.05)
    s = compare_runs(jsb, oracle, ("lattice", 2, 20, 20), p, reps=2)
    assert max(int(x["n_clones"]) for x in s) > 100 and all(int(x["status"]) == 0 for x in s)


def test_arena_compaction_with_tiered_lists(jsb, oracle, monkeypatch):
    # lists longer than one tiered-vector segment (256 entries) in an arena so small that it is compacted
    # again and again: tiered -> flat squeeze, aligned re-spread, head resets, growth of tiered blocks,
    # the shrink-to-one-segment rotation (voter turnover empties and refills clones)
    monkeypatch.setenv("JSB_DEBUG_ARENA_ENTRIES", "16000")
    p = dict(DEFAULT, rule="voter", Tf=90.0, rate_birth=0.5, mu_driver=0.003, rate_death=0.05, rate_migration=0.05)
    s = compare_runs(jsb, oracle, ("lattice", 2, 64, 64), p, reps=4)
    assert max(int(x["n_alive"]) for x in s) > 1000 and all(int(x["status"]) == 0 for x in s)
    # and with excisions in the middle of it (tiered lists rebuilt flat through the scratch array)
    p = dict(DEFAULT, Tf=70.0, rate_birth=0.6, rate_death=0.03, t_bottleneck=[30.0, 50.0], ratio_bottleneck=[0.6, 0.3],
             quirk_flags=1)
    s = compare_runs(jsb, oracle, ("lattice", 2, 64, 64), p, reps=4)
    assert max(int(x["n_excised"]) for x in s) > 300 and all(int(x["status"]) == 0 for x in s)


def test_clone_capacity_stop(jsb, oracle):
    # 1000 clones = every label of Set(1:1000) used: the reference throws at :371; both sides stop
    p = dict(DEFAULT, rule="voter", Tf=150.0, mu_driver=0.1, rate_death=0.05)
    s = compare_runs(jsb, oracle, ("lattice", 2, 20, 20), p, reps=1)
    assert int(s[0]["status"]) == 2 and int(s[0]["n_clones"]) == 1000


def test_gpu_matches_committed_golden_vectors(jsb, oracle):
    # fixtures made by tests/golden/make_golden.py (CPU oracle); the GPU must reproduce the files
    from golden.make_golden import CASES
    from jspace_b200 import _lib as L
    here = os.path.dirname(__file__)
    for name, (spec, params, seed, stream) in CASES.items():
        ref = np.load(os.path.join(here, "golden", f"{name}.npz"))
        g = (jsb.Graph.lattice(spec[2], spec[3], spec[1]) if spec[0] == "lattice"
             else jsb.Graph.dense_file(os.path.join(here, "golden", spec[1])))
        res = jsb.run_ensemble(g, jsb.Point(**params), stream + 1, seed=seed, g_begin=stream, g_end=stream + 1,
                               flags=L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS)
        assert res.summaries[0].tobytes() == ref["summary"][0].tobytes(), name
        assert res.events(0).tobytes() == ref["events"].tobytes(), name
        assert res.clones(0).tobytes() == ref["clones"].tobytes(), name
        c, i = res.lattice(0)
        assert np.array_equal(c, ref["clone_of_site"]) and np.array_equal(i, ref["cell_of_site"]), name
        assert np.array_equal(res.list(0, 0), ref["cs_alive"]), name
        res.close()


def test_python_mirror_of_reference_api(jsb, oracle, tmp_path):
    # simulate_evolution / Start_J_Space keep the reference's signatures and 7-tuple
    G = jsb.spatial_graph(30, 30, jsb.PhiloxSeed(1234), dim=2, n_cell=1)
    seed = jsb.PhiloxSeed(1234)
    df, G2, n_alive, set_mut, gs, ca, alpha = jsb.simulate_evolution(
        G, 60.0, 0.2, 0.01, 0.01, 0.05, 0.4, 0.1, "contact", seed, Time_of_sampling=[10, 20],
        t_bottleneck=[30.0], ratio_bottleneck=[0.5])
    o = oracle.Run(oracle.Graph.lattice(2, 30, 30), oracle.make_params(
        Tf=60.0, mu_driver=0.05, t_bottleneck=[30.0], ratio_bottleneck=[0.5], t_snapshot=[10.0, 20.0]), 1234, 0,
        faithful=False)
    assert df.records.tobytes() == o.events.tobytes()
    assert n_alive == o.series.tolist() and alpha == o.clones["alpha"].tolist()
    assert set_mut[0] == 1 and all(g[0] == 1 for g in set_mut[1:]) and len(gs) == 2
    assert sorted(G2.cells_alive()) == sorted(o.lists[0].tolist())
    assert len(gs[1].cells_alive()) == int(o.snapshots["n_alive"][1])
    v = G2.cells_alive()[0]
    assert G2.get_prop(v, "Fit") == alpha[G2.get_prop(v, "Subpop") - 1]
    # the next call on the same seed object uses the next stream, like a shared MersenneTwister
    df2 = jsb.simulate_evolution(G, 30.0, 0.2, 0.01, 0.01, 0.05, 0.4, 0.1, "contact", seed)[0]
    o2 = oracle.Run(oracle.Graph.lattice(2, 30, 30), oracle.make_params(Tf=30.0, mu_driver=0.05), 1234, 1, faithful=False)
    assert df2.records.tobytes() == o2.events.tobytes()
    # Start_J_Space: same TOML keys as the reference's Parameters.toml / Config.toml
    (tmp_path / "P.toml").write_text(
        '[[Graph]]\nrow = 20\ncol = 20\ndim = 2\nN_starting_cells = 1\n[[Dynamic]]\nModel = "contact"\nMax_time = 50.0\n'
        'rate_birth = 0.2\nrate_death = 0.01\nrate_migration = 0.01\ndrive_mut_rate = 0.01\n'
        'average_driver_mut_rate = 0.4\nstd_driver_mut_rate = 0.1\nt_bottleneck = [25.0]\nratio_bottleneck = [0.5]\n')
    (tmp_path / "C.toml").write_text(
        f'[[Config]]\nseed = 1234\npath_to_save_files = "{tmp_path}/out/"\nGenerate_graph = 1\n'
        'Path_to_Graph = ""\nTree_Driver_Configure = 0\nedgelist_treedriver = ""\ndriver_birth_rates = ""\n'
        '[[OutputGT]]\nDriver_list = 1\n[[FileOutputPlot]]\nTime_of_sampling = [10, 20]\nGraph_configuration = 1\n')
    out = jsb.Start_J_Space(str(tmp_path / "P.toml"), str(tmp_path / "C.toml"))
    files = sorted(os.listdir(tmp_path / "out"))
    assert files == ["CA_subpop.csv", "Dinamica.csv", "DriverList.csv", "n_cell_alive.csv"]
    assert (tmp_path / "out" / "DriverList.csv").read_text().splitlines()[0] == "1,0.2"
    assert len((tmp_path / "out" / "n_cell_alive.csv").read_text().splitlines()) == len(out[2])
    assert (tmp_path / "out" / "Dinamica.csv").read_text().splitlines()[0] == "Event,Time,Cell,Notes"


def test_genealogy_of_sampled_cells(jsb, oracle):
    # jsb_result_genealogy == create_matrix_relational (reference :60-104) restated on the oracle's log
    from jspace_b200 import _lib as L
    g = jsb.Graph.lattice(30, 30, 2)
    p = dict(DEFAULT, Tf=60.0, mu_driver=0.05)
    res = jsb.run_ensemble(g, jsb.Point(**p), 1, seed=5, flags=L.OUT_EVENTS | L.OUT_LATTICE)
    ev, (clone, cid), clones = res.events(0), res.lattice(0), res.clones(0)
    alive = cid[cid > 0]
    rng = np.random.default_rng(0)
    sample = rng.choice(alive, size=min(10, len(alive)), replace=False).astype(np.uint32)
    rel = res.genealogy(0, sample)
    # restatement of the backward scan
    wanted = set(int(x) for x in sample)
    rows = []
    for e in ev[::-1]:
        if e["type"] not in (L.EV_DUPLICATE, L.EV_MUTATION) or int(e["cell"]) not in wanted:
            continue
        wanted.discard(int(e["cell"]))
        wanted.add(int(e["parent"]))
        sub = int(clones["parent"][e["clone"] - 1]) if e["type"] == L.EV_MUTATION else int(e["clone"])
        rows.append((int(e["parent"]), int(e["cell"]), float(e["time"]), sub, int(e["type"])))
    got = [(int(r["father"]), int(r["child"]), float(r["time"]), int(r["subpop_child"]), int(r["type"])) for r in rel]
    assert got == rows and len(rows) >= len(sample)
    assert wanted == {1}  # every sampled lineage coalesces in the seed cell
    res.close()


def test_log_pool_exhaustion_keeps_a_prefix_of_the_log(jsb, monkeypatch):
    # An event-log pool that is too small: a replicate that cannot claim another chunk keeps the rows it has (whole
    # chunks of 1024), carries on without a log and reports JSB_ST_LOG_OVERFLOW — also when it was parked by the pilot
    # pass and resumed, and also when the outputs are streamed (second call).  Everything else must equal the run with
    # a pool that is large enough.
    from jspace_b200 import _lib as L
    monkeypatch.setenv("JSB_DEBUG_GRID", "3")
    monkeypatch.setenv("JSB_PILOT_ITERS", "2000")
    g = jsb.Graph.lattice(64, 64, 2)
    p = jsb.Point(**dict(DEFAULT, Tf=90.0, rate_birth=0.5))
    flags = L.OUT_EVENTS | L.OUT_LATTICE
    full = jsb.run_ensemble(g, p, 40, seed=3, flags=flags, warps_per_sm=2, log_pool_bytes=256 << 20)
    assert np.all(full.summaries["status"] == L.ST_OK) and int(full.summaries["n_events"].max()) > 4096
    needed = int(np.ceil(full.summaries["n_events"] / 1024.0).sum())  # chunks of 1024 rows
    for attempt in range(2):  # the second call finds a pinned block and a gather buffer: streamed
        small = jsb.run_ensemble(g, p, 40, seed=3, flags=flags, warps_per_sm=2, log_pool_bytes=(needed * 2 // 3) * 1024 * 32)
        st = small.summaries["status"]
        assert int((st == L.ST_LOG_OVERFLOW).sum()) >= 2 and int((st == L.ST_OK).sum()) >= 2, st
        a, b = small.summaries.copy(), full.summaries.copy()
        a["status"] = 0
        b["status"] = 0
        assert a.tobytes() == b.tobytes()  # the trajectories themselves are untouched
        for i in range(40):
            es, ef = small.events(i), full.events(i)
            if st[i] == L.ST_LOG_OVERFLOW:
                assert len(es) < len(ef) and len(es) % 1024 == 0
            else:
                assert len(es) == len(ef)
            assert es.tobytes() == ef[:len(es)].tobytes(), i
            assert small.lattice(i)[0].tobytes() == full.lattice(i)[0].tobytes()
        small.close()
    full.close()


@pytest.mark.parametrize("rule,method", [("contact", 1), ("voter", 1), ("hvoter", 2)])
def test_device_side_series_equals_oracle_series_and_log_replay(jsb, oracle, rule, method):
    # JSB_OUT_SERIES: list_len_node_occ / CA_Alive_TOT reduced from the log on the device (no log on the host):
    # against the oracle's own series, and against the host replay of the log of the same run
    from jspace_b200 import _lib as L
    from jspace_b200.api import replay_series
    from parity import make_graphs
    p = dict(DEFAULT, Tf=70.0, rule=rule, t_bottleneck=[25.0, 45.0], ratio_bottleneck=[0.6, 0.3], t_snapshot=[10.0, 30.0])
    if method == 2:
        p = dict(method=2, Tf=70.0, rate_death=0.02, rate_migration=0.01, mu_driver=0.05, rule=rule, tree_edges=TREE_EDGES,
                 node_rate=TREE_RATES, root_node=0, root_rate=0.3, t_bottleneck=[30.0], ratio_bottleneck=[0.5],
                 t_snapshot=[20.0])
    gj, go = make_graphs(jsb, oracle, ("lattice", 2, 30, 30))
    only_series = jsb.run_ensemble(gj, jsb.Point(**p), 24, seed=11, flags=L.OUT_SERIES)
    both = jsb.run_ensemble(gj, jsb.Point(**p), 24, seed=11, flags=L.OUT_SERIES | L.OUT_EVENTS)
    assert only_series.summaries.tobytes() == both.summaries.tobytes()
    with pytest.raises(L.JSpaceError):
        only_series.events(0)
    exact = 0
    for i in range(24):
        o = oracle.Run(go, oracle.make_params(**p), 11, i)
        n_occ, ca_tot = only_series.series(i)
        # (an excision that removes no cell pushes in the reference but writes no log row, DESIGN.md 6: the
        # device-side series then has one entry less than n_effective + 1)
        assert len(n_occ) - 1 <= int(only_series.summaries["n_effective"][i])
        if len(n_occ) - 1 == int(only_series.summaries["n_effective"][i]):
            assert n_occ == o.series.tolist(), i
            exact += 1
        assert only_series.series_rows(i)[0].tobytes() == both.series_rows(i)[0].tobytes()
        ev = both.events(i)
        n_occ2, ca_tot2 = replay_series(ev, 1, [int(s["n_events_before"]) for s in both.snapshots(i)])
        assert ca_tot == ca_tot2 and n_occ == n_occ2
        if len(ca_tot):
            assert ca_tot[-1] == o.clones["count"].tolist()
    assert exact >= 20
    only_series.close()
    both.close()


def test_output_flag_combinations_agree(jsb, monkeypatch):
    # every combination of output flags goes through a different mix of paths (pilot parked/resumed, streamed or
    # gathered outputs, series reduction, slots by finish rank or by index): the outputs they share must be identical,
    # on a first round of calls and on a second one that finds the cached host / device buffers
    from jspace_b200 import _lib as L
    monkeypatch.setenv("JSB_DEBUG_GRID", "4")
    monkeypatch.setenv("JSB_PILOT_ITERS", "500")
    g = jsb.Graph.lattice(50, 50, 2)
    p = jsb.Point(Tf=80.0, rate_birth=0.3, rate_death=0.02, rate_migration=0.01, mu_driver=0.02, adv_mean=0.4, adv_std=0.1,
                  rule="contact", t_bottleneck=[30.0], ratio_bottleneck=[0.5])
    ref = None
    for rnd in range(2):
        for flags in (L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS, L.OUT_LATTICE, L.OUT_LISTS, L.OUT_LATTICE | L.OUT_LISTS,
                      L.OUT_SERIES | L.OUT_LATTICE, L.OUT_EVENTS, 0):
            r = jsb.run_ensemble(g, p, 64, seed=9, flags=flags, warps_per_sm=2)
            lat = [r.lattice(i)[0].tobytes() + r.lattice(i)[1].tobytes() for i in range(64)] if flags & L.OUT_LATTICE else None
            lis = [r.list(i, 0).tobytes() + r.list(i, 1).tobytes() for i in range(64)] if flags & L.OUT_LISTS else None
            evs = [r.events(i).tobytes() for i in range(64)] if flags & L.OUT_EVENTS else None
            if ref is None:
                ref = dict(sm=r.summaries.tobytes(), lat=lat, lis=lis, evs=evs)
            assert r.summaries.tobytes() == ref["sm"], (rnd, flags)
            assert lat is None or lat == ref["lat"], (rnd, flags)
            assert lis is None or lis == ref["lis"], (rnd, flags)
            assert evs is None or evs == ref["evs"], (rnd, flags)
            r.close()
