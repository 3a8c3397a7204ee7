"""Rejection-free mode (JSB_MODE_REJECTION_FREE): same stochastic law as the
reference loop, not the same trajectories.  Checked by (a) two-sample KS tests (alpha = 0.01) of
ensemble statistics against the CPU oracle of the reference algorithm, per contact rule / method /
excision, and (b) structural invariants: the event log replays to the final lattice, counts add
up, ids are unique."""
import numpy as np
import pytest
from scipy import stats

from parity import TREE_EDGES, TREE_RATES
from test_fullsize_gpu import replay_lattice

pytestmark = pytest.mark.gpu

BASE = dict(Tf=80.0, rate_birth=0.25, rate_death=0.02, rate_migration=0.02, mu_driver=0.03, adv_mean=0.4, adv_std=0.1)
CASES = {
    "contact": dict(BASE, rule="contact"),
    "contact_excision": dict(BASE, rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.6]),
    "voter": dict(BASE, rule="voter", Tf=50.0, mu_driver=0.005),
    "hvoter": dict(BASE, rule="hvoter", Tf=50.0, mu_driver=0.005),
    "tree": dict(method=2, Tf=80.0, rate_death=0.02, rate_migration=0.02, mu_driver=0.05, rule="contact",
                 tree_edges=TREE_EDGES, node_rate=TREE_RATES, root_node=0, root_rate=0.25),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_fast_mode_matches_reference_law(jsb, oracle, case):
    from jspace_b200 import _lib as L
    params = CASES[case]
    n = 400
    g = jsb.Graph.lattice(32, 32, 2)
    res = jsb.run_ensemble(g, jsb.Point(**params), n, seed=777, flags=L.MODE_REJECTION_FREE)
    fs = res.summaries.copy()
    frac_fast = [res.clones(i)["count"].max() / fs["n_alive"][i] for i in range(n) if fs["n_alive"][i] > 0]
    res.close()
    assert np.all(fs["status"] == 0)
    go = oracle.Graph.lattice(2, 32, 32)
    runs = [oracle.Run(go, oracle.make_params(**params), 4242, 5000 + i, faithful=False, want=("clones",))
            for i in range(n)]
    osum = np.array([r.summary for r in runs])
    frac_ref = [r.clones["count"].max() / r.summary["n_alive"] for r in runs if r.summary["n_alive"] > 0]
    for field in ("n_alive", "n_clones", "n_births", "n_deaths", "n_migrations", "n_displaced", "n_excised"):
        a, b = fs[field].astype(float), osum[field].astype(float)
        if a.max() == a.min() == b.max() == b.min():
            continue
        p = stats.ks_2samp(a, b).pvalue
        assert p > 0.01, (case, field, p, a.mean(), b.mean())
    assert stats.ks_2samp(frac_fast, frac_ref).pvalue > 0.01
    # extinction probability
    ext_f, ext_r = (fs["n_alive"] == 0).mean(), (osum["n_alive"] == 0).mean()
    assert abs(ext_f - ext_r) < 4 * np.sqrt(ext_r * (1 - ext_r) / n + 1e-9) + 0.02


@pytest.mark.parametrize("rule", ["contact", "voter", "hvoter"])
def test_fast_mode_incremental_tree_is_consistent(jsb, rule):
    # JSB_QUIRK_DEBUG_VERIFY_TREE: after every event the incrementally updated rates around the
    # changed sites must equal the rates recomputed from scratch (else status 5)
    from graphs import random_geometric_csr
    from jspace_b200 import _lib as L
    p = jsb.Point(Tf=60.0, rate_birth=0.3, rate_death=0.03, rate_migration=0.05, mu_driver=0.01, rule=rule,
                  t_bottleneck=[30.0], ratio_bottleneck=[0.4], quirk_flags=L.QUIRK_DEBUG_VERIFY_TREE)
    for g in (jsb.Graph.lattice(24, 24, 2), jsb.Graph.lattice(27, 27, 3)):
        res = jsb.run_ensemble(g, p, 24, seed=9, flags=L.MODE_REJECTION_FREE)
        assert np.all(res.summaries["status"] == 0), np.bincount(res.summaries["status"])
        assert res.summaries["n_births"].max() > 300
        res.close()
    rowptr, colidx = random_geometric_csr(1500, seed=5)
    g = jsb.Graph.csr(rowptr, colidx)
    g.set_seed_candidates([7])
    res = jsb.run_ensemble(g, p, 24, seed=9, flags=L.MODE_REJECTION_FREE)
    assert np.all(res.summaries["status"] == 0), np.bincount(res.summaries["status"])
    res.close()


def test_fast_mode_invariants_and_log_replay(jsb):
    from jspace_b200 import _lib as L
    g = jsb.Graph.lattice(60, 60, 2)
    p = jsb.Point(Tf=120.0, mu_driver=0.003, rule="voter", t_bottleneck=[60.0], ratio_bottleneck=[0.5],
                  t_snapshot=[30.0, 90.0])
    res = jsb.run_ensemble(g, p, 16, seed=5, flags=L.MODE_REJECTION_FREE | L.OUT_EVENTS | L.OUT_LATTICE)
    s = res.summaries
    assert np.all(s["status"] == 0) and s["n_births"].max() > 1000
    seed_site = int(g.seed_candidates()[0])
    for i in range(res.count):
        ev, cl = res.events(i), res.clones(i)
        clone, cid = res.lattice(i)
        rc, ri = replay_lattice(g.nv, ev, [(seed_site, 1)])
        assert np.array_equal(rc, clone) and np.array_equal(ri, cid)
        assert int(np.count_nonzero(clone)) == s["n_alive"][i] == int(cl["count"].sum())
        assert np.array_equal(np.bincount(clone, minlength=len(cl) + 1)[1:], cl["count"])
        born = ev[ev["type"] <= 1]
        assert len(np.unique(born["cell"])) == len(born) == 2 * s["n_births"][i]
        t = ev["time"][(ev["flags"] & 4) == 0]
        assert np.all(np.diff(t) >= 0) and (len(t) == 0 or t[-1] < 120.0 + 5.0)
        assert s["n_iterations"][i] == s["n_births"][i] + s["n_deaths"][i] + s["n_migrations"][i]
        if s["n_alive"][i] > 0:
            assert len(res.snapshots(i)) == 2
    with pytest.raises(jsb.JSpaceError, match="JSB_OUT_LISTS"):
        jsb.run_ensemble(g, p, 1, flags=L.MODE_REJECTION_FREE | L.OUT_LISTS)
    res.close()


def test_fast_mode_graph_and_3d(jsb, oracle):
    from graphs import random_geometric_csr
    from jspace_b200 import _lib as L
    rowptr, colidx = random_geometric_csr(2500, seed=11)
    params = dict(BASE, rule="contact", Tf=70.0)
    n = 300
    g = jsb.Graph.csr(rowptr, colidx)
    g.set_seed_candidates([40])
    res = jsb.run_ensemble(g, jsb.Point(**params), n, seed=3, flags=L.MODE_REJECTION_FREE)
    go = oracle.Graph.csr(rowptr, colidx)
    osum = np.array([oracle.Run(go, oracle.make_params(**params), 9, 700 + i, candidates=[40], faithful=False,
                                want=()).summary for i in range(n)])
    for field in ("n_alive", "n_births", "n_clones"):
        assert stats.ks_2samp(res.summaries[field].astype(float), osum[field].astype(float)).pvalue > 0.01, field
    res.close()
    g3 = jsb.Graph.lattice(30, 30, 3)
    r3 = jsb.run_ensemble(g3, jsb.Point(**dict(params, rule="hvoter", mu_driver=0.003)), 64, seed=3, flags=L.MODE_REJECTION_FREE)
    assert np.all(r3.summaries["status"] == 0) and r3.summaries["n_alive"].max() > 100
    r3.close()
