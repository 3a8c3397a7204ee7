"""The CPU oracle against what the reference pins for this path.  The reference has no tests or
golden vectors; what exists is (a) the shipped fileoutput/DriverList.csv, usable as
a format + statistical fixture, (b) the structural invariants of the algorithm, (c) committed
oracle outputs (tests/golden/*.npz, made by tests/golden/make_golden.py) as a regression pin."""
import os

import numpy as np
import pytest
from scipy import stats

HERE = os.path.dirname(__file__)
DEFAULT = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4,
               adv_std=0.1, rule="contact")


def parse_driver_list(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith('"'):
                g, a = line[1:].split('",')
                geno = [int(x) for x in g.strip("[]").split(",")]
            else:
                g, a = line.split(",")
                geno = [int(g)]
            rows.append((geno, float(a)))
    return rows


def increments(rows):
    by_geno = {tuple(g): a for g, a in rows}
    return np.array([a - by_geno[tuple(g[:-1])] for g, a in rows if len(g) > 1])


def test_reference_driver_list_fixture_is_consistent():
    rows = parse_driver_list(os.path.join(HERE, "golden", "reference_DriverList.csv"))
    assert len(rows) == 74 and rows[0] == ([1], 0.2)
    labels = [g[-1] for g, _ in rows[1:]]
    assert len(set(labels)) == len(labels) and all(2 <= x <= 1000 for x in labels)
    seen = {(1,)}
    for g, _ in rows[1:]:
        assert tuple(g[:-1]) in seen  # prefix closed, parent listed before child
        seen.add(tuple(g))
    inc = increments(rows)
    assert abs(inc.mean() - 0.3935) < 1e-3 and inc.min() > 0


def test_oracle_driver_increments_match_reference_distribution(oracle):
    """alpha[child] - alpha[parent] = |N(0.4, 0.1)| (src/J_Space.jl:565-566): the oracle's clone
    tables must be statistically indistinguishable from the reference's shipped driver list."""
    ref = increments(parse_driver_list(os.path.join(HERE, "golden", "reference_DriverList.csv")))
    g = oracle.Graph.lattice(2, 100, 100)
    mine = []
    for s in range(3):
        r = oracle.Run(g, oracle.make_params(**DEFAULT), 1234, s, faithful=False, want=("clones",))
        c = r.clones
        mine.extend(c["alpha"][i] - c["alpha"][c["parent"][i] - 1] for i in range(1, len(c)))
    mine = np.array(mine)
    assert len(mine) > 100
    assert stats.ks_2samp(ref, mine).pvalue > 0.01
    folded = stats.foldnorm(c=0.4 / 0.1, scale=0.1)
    assert stats.kstest(mine, folded.cdf).pvalue > 0.01
    # labels: unique, in 2..1000
    assert len(set(c["label"].tolist())) == len(c) and c["label"][0] == 1


@pytest.mark.parametrize("rule", ["contact", "voter", "hvoter"])
def test_oracle_invariants(oracle, rule):
    g = oracle.Graph.lattice(2, 30, 30)
    p = dict(DEFAULT, Tf=60.0, rule=rule, mu_driver=0.03, t_bottleneck=[30.0], ratio_bottleneck=[0.5])
    for s in range(3):
        r = oracle.Run(g, oracle.make_params(**p), 99, s, faithful=True)
        ev, cl, sm = r.events, r.clones, r.summary
        alive_sites = np.nonzero(r.clone_of_site)[0] + 1
        # occupancy <-> ordered lists
        assert sorted(r.lists[0].tolist()) == alive_sites.tolist() and len(alive_sites) == sm["n_alive"]
        assert sum(len(x) for x in r.lists[1:]) == sm["n_alive"]
        for ci, lst in enumerate(r.lists[1:], start=1):
            assert all(r.clone_of_site[v - 1] == ci for v in lst) and cl["count"][ci - 1] == len(lst)
        # ids unique; every non-seed id is created by exactly one Duplicate/Mutation row
        born = ev[(ev["type"] == 0) | (ev["type"] == 1)]
        assert len(set(born["cell"].tolist())) == len(born)
        ids = r.cell_of_site[r.cell_of_site > 0]
        assert len(set(ids.tolist())) == len(ids) and ids.max() < sm["next_cell_id"]
        # genealogy is a forest rooted at the seed: parents are created before children
        created = {1: -1}
        for i, e in enumerate(born):
            assert int(e["parent"]) in created
            created[int(e["cell"])] = i
        # times are non-decreasing except excision rows (logged at t_bottleneck)
        t = ev["time"][(ev["flags"] & 4) == 0]
        assert np.all(np.diff(t) >= 0)
        # the series: one entry per effective event, ends at n_alive
        assert len(r.series) == sm["n_effective"] + 1 and r.series[-1] == sm["n_alive"]
        # clone table: parent precedes child, alpha of a child >= alpha of its parent
        for i in range(1, len(cl)):
            assert 1 <= cl["parent"][i] <= i and cl["alpha"][i] >= cl["alpha"][cl["parent"][i] - 1]


def test_faithful_and_indexed_membership_agree(oracle):
    g = oracle.Graph.lattice(2, 40, 40)
    p = oracle.make_params(**dict(DEFAULT, Tf=80.0))
    a = oracle.Run(g, p, 5, 0, faithful=True)
    b = oracle.Run(g, p, 5, 0, faithful=False)
    assert a.events.tobytes() == b.events.tobytes() and a.summary == b.summary


def test_golden_vectors_reproduce(oracle):
    from golden.make_golden import CASES, run_case
    for name in CASES:
        got = run_case(oracle, name)
        ref = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
        for k in ref.files:
            assert got[k].tobytes() == ref[k].tobytes(), (name, k)


def test_lattice_graph_restatement(oracle):
    # Graphs.grid numbering: first dimension fastest, neighbours ascending; J_Space.jl:50-54,131-133
    g = oracle.Graph.lattice(2, 4, 3)
    assert g.nv == 12 and g.neighbors(1).tolist() == [2, 5] and g.neighbors(6).tolist() == [2, 5, 7, 10]
    g3 = oracle.Graph.lattice(3, 350, 350)
    assert g3.nv == 50 ** 3 and g3.neighbors(61426).tolist() == [58926, 61376, 61425, 61427, 61476, 63926]
    ga = oracle.Graph.dense_file(os.path.join(HERE, "golden", "Example_adj_matrix.txt"))
    assert ga.nv == 10 and ga.closeness_argmax().tolist() == [1, 2]


@pytest.mark.parametrize("model,extra", [("contact", dict(t_bottleneck=[25.0], ratio_bottleneck=[0.8])),
                                         ("contact", dict(t_bottleneck=[30.0], ratio_bottleneck=[0.03])),
                                         ("voter", {}), ("hvoter", {})])
def test_oracle_equals_independent_python_restatement(oracle, model, extra):
    """tests/pymodel.py restates the reference loop a second time, in plain Python, from the
    reference source and the draw contract; the C++ oracle must reproduce it bit for bit."""
    import pymodel
    base = dict(Tf=45.0, rate_birth=0.25, rate_death=0.03, rate_migration=0.03, mu_driver=0.06, adv_mean=0.4, adv_std=0.1)
    for (row, col), stream in (((9, 7), 0), ((12, 12), 3), ((6, 10), 5)):
        df, alpha, set_mut, it, t = pymodel.simulate(row, col, base["Tf"], base["rate_birth"], base["rate_death"],
                                                     base["rate_migration"], base["mu_driver"], base["adv_mean"],
                                                     base["adv_std"], model, 4321, stream,
                                                     extra.get("t_bottleneck", ()), extra.get("ratio_bottleneck", ()))
        r = oracle.Run(oracle.Graph.lattice(2, row, col), oracle.make_params(**dict(base, rule=model, **extra)), 4321,
                       stream, faithful=True)
        ev = r.events
        assert len(ev) == len(df) and r.summary["n_iterations"] == it and r.summary["t_final"] == t
        for k, row_py in enumerate(df):
            e = ev[k]
            got = (int(e["type"]), float(e["time"]), int(e["cell"]), int(e["parent"]), int(e["site"]), int(e["site2"]),
                   int(e["clone"]), int(e["label"]), int(e["flags"]))
            assert got == row_py, (model, (row, col), k, got, row_py)
        assert r.clones["alpha"].tolist() == alpha
        assert [(int(c["parent"]), int(c["label"])) for c in r.clones] == set_mut
