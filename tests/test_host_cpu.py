"""Host-side mirror of the reference interface (no GPU needed): driver-tree files, series replay,
CSV writers in the reference's formats, ensemble sharding."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(__file__)


def test_read_driver_tree_matches_reference_files(jsb, tmp_path):
    from jspace_b200.api import read_driver_tree
    e = tmp_path / "tree.txt"
    a = tmp_path / "adv.txt"
    e.write_text("driver_1 driver_2\ndriver_2 driver_3\ndriver_3 driver_4")   # utility/Tree_driver_example.txt
    a.write_text("driver_1 0.2\ndriver_2 0.5\ndriver_3 0.55\ndriver_4 0.6 ")   # utility/driver_advantage.txt
    names, edges, rate, root, root_rate = read_driver_tree(str(e), str(a))
    assert names == ["driver_1", "driver_2", "driver_3", "driver_4"]
    assert edges == [(0, 1), (1, 2), (2, 3)] and rate == [0.2, 0.5, 0.55, 0.6] and root == 0 and root_rate == 0.2
    a.write_text("driver_2 0.5\ndriver_1 0.2\ndriver_3 0.55\ndriver_4 0.6")
    # alpha[1] is the rate in the FIRST row even if that row is not the root (src/J_Space.jl:881)
    assert read_driver_tree(str(e), str(a))[4] == 0.5
    a.write_text("driver_1 0.2\n")
    with pytest.raises(ValueError, match="no birth rate"):
        read_driver_tree(str(e), str(a))


def test_series_replay_matches_oracle(jsb, oracle):
    from jspace_b200.api import replay_series
    base = dict(Tf=70.0, rate_birth=0.2, rate_death=0.02, rate_migration=0.02, mu_driver=0.05, adv_mean=0.4,
                adv_std=0.1)
    for rule, extra in (("contact", dict(t_bottleneck=[40.0], ratio_bottleneck=[0.6])), ("voter", {}), ("hvoter", {})):
        g = oracle.Graph.lattice(2, 28, 28)
        r = oracle.Run(g, oracle.make_params(**dict(base, rule=rule, **extra)), 11, 0, faithful=False)
        series, ca_tot = replay_series(r.events, 1)
        assert series == r.series.tolist()
        assert ca_tot[-1] == r.clones["count"].tolist() and len(ca_tot) == len(series) - 1


def test_genotypes_and_csv_formats(jsb, tmp_path):
    from jspace_b200 import _lib as L
    from jspace_b200.api import genotypes_from_clones
    from jspace_b200.start import julia_float, write_driver_list
    clones = np.zeros(4, dtype=L.CLONE_DTYPE)
    clones["label"] = [1, 143, 3, 431]
    clones["parent"] = [0, 1, 1, 2]
    clones["alpha"] = [0.2, 0.6253257383233928, 0.4340527798779604, 1.0e-5]
    g = genotypes_from_clones(clones)
    assert g == [1, [1, 143], [1, 3], [1, 143, 431]]
    assert genotypes_from_clones(clones[:2], ["driver_1"] * 143) == ["driver_1", ["driver_1", "driver_1"]]
    p = tmp_path / "DriverList.csv"
    write_driver_list(str(p), g, clones["alpha"], header=False)
    lines = p.read_text().splitlines()
    # first rows of the reference's shipped fileoutput/DriverList.csv
    ref = open(os.path.join(HERE, "golden", "reference_DriverList.csv")).read().splitlines()
    assert lines[0] == ref[0] == "1,0.2" and lines[1] == ref[1] == '"[1, 143]",0.6253257383233928'
    assert lines[2] == ref[2] and lines[3] == '"[1, 143, 431]",1.0e-5'
    assert julia_float(200.0) == "200.0" and julia_float(1e16) == "1.0e16" and julia_float(0.0001) == "0.0001"


def test_shard_ranges_partition_the_ensemble(jsb):
    for total in (0, 1, 7, 4096, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            r = [jsb.shard_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_series_rows_to_reference_series():
    # jsb_series_row -> list_len_node_occ / CA_Alive_TOT (host side of JSB_OUT_SERIES); the device side is compared with
    # the oracle in tests/test_parity_gpu.py
    import numpy as np
    from jspace_b200 import _lib as L
    from jspace_b200.ensemble import series_from_rows
    M = L.SERIES_MORE
    rows = np.array([(2, 1, 0),          # birth in clone 1
                     (3, 2, 0),          # Mutation: clone 2 appears with one cell
                     (3, 0, 0),          # migration
                     (2, 0, 1),          # death in clone 1
                     (2, 2, 1),          # voter birth of clone 2 displacing a cell of clone 1
                     (1 | M, 0, 2), (0, 0, 2)],  # excision removing two cells of clone 2: one push
                    dtype=L.SERIES_DTYPE)
    n_occ, ca = series_from_rows(rows, rows_before_snapshot=[2, 7], n0=1)
    assert n_occ == [1, 2, 3, 3, 2, 2, 0]
    assert ca == [[2], [2, 1], [2, 1], [2, 1], [1, 1], [0, 2], [0, 0], [0, 0]]
    assert series_from_rows(rows[:0], [0], n0=1) == ([1], [[1]])
