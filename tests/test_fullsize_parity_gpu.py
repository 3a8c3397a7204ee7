"""BASELINE.json's configurations at FULL size, CUDA path (through the C ABI) against the CPU oracle,
bit for bit: every log row, event time, fitness value, final lattice and ordered list.

With `faithful=False` (O(1) occupancy test instead of the reference's O(n) `pos in cs_alive` scan — same
draws, same decisions) the oracle finishes a 2.4-5.4 M-iteration 200x200 replicate in 4-14 s, so the
headline configuration is compared directly, inside real ensembles, which is where the scheduling
machinery of the kernel is live: lookahead helper warps, the decoupled clock (>= 64 clones), arena
compaction, the pilot pass with its longest-first order, static dealing and the dynamic work counter.

  C2  200x200, Parameters.toml dynamics + one excision, seed 1234: streams {0, 6, 7, 13} of a 32-replicate
      ensemble (every replicate owns an SM and its helpers), and streams {0, 1, 3, 4, 5, 7} of the full
      4096-replicate ensemble (pilot + order + counter, exactly what bench.py times)
  C3  50^3 lattice (row = col = 350, dim 3: u32 site ids), linear 4-driver tree, excision at t = 100,
      `contact` and `hvoter`
  C4  ~100 k-node random geometric graph through the CSR path
  C5  the corners of the birth/death/advantage sweep grid on 100x100
  and the pilot/order/counter path on a small lattice with a forced tiny grid, with and without pilot.
Reference: src/J_Space.jl:638-842 (method 1), :848-1104 (method 2)."""
import numpy as np
import pytest

from parity import TREE_EDGES, TREE_RATES, compare_runs

pytestmark = pytest.mark.gpu

C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4,
          adv_std=0.1, rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])


def test_c2_heaviest_streams_in_a_32_replicate_ensemble(jsb, oracle):
    # 32 replicates on 148 SMs: static dealing, up to 1 + kMaxHelp warps per replicate, decoupled clock
    s = compare_runs(jsb, oracle, ("lattice", 2, 200, 200), C2, seed=1234, reps=32, only=[0, 6, 7, 13])
    assert min(int(x["n_iterations"]) for x in s) > 2_000_000
    assert min(int(x["n_clones"]) for x in s) > 300


def test_c2_members_of_the_full_4096_replicate_ensemble(jsb, oracle):
    # the default path of the headline configuration: pilot pass, order[] remap, boustrophedon static
    # dealing, dynamic counter starting at `slots`, helpers attaching in the tail
    got = []
    s = compare_runs(jsb, oracle, ("lattice", 2, 200, 200), C2, seed=1234, reps=4096, only=[0, 1, 3, 4, 5, 7],
                     check_lists=False, log_pool_bytes=14 << 30, summaries_out=got)
    assert max(int(x["n_iterations"]) for x in s) > 2_000_000
    from jspace_b200 import _lib as L
    assert np.all(got[0]["status"] == L.ST_OK)


def test_pilot_order_and_counter_path_vs_oracle(jsb, oracle, monkeypatch):
    # more replicates than warp slots (4 blocks x 2 warps): pilot + longest-first order + counter, every
    # replicate against the oracle; and the same ensemble without the pilot must give identical summaries
    p = dict(C2, Tf=70.0, t_bottleneck=[30.0])
    monkeypatch.setenv("JSB_DEBUG_GRID", "4")
    monkeypatch.setenv("JSB_PILOT_ITERS", "300")
    a, b = [], []
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, seed=77, reps=48, warps_per_sm=2, summaries_out=a)
    monkeypatch.setenv("JSB_NO_PILOT", "1")
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, seed=77, reps=48, warps_per_sm=2, summaries_out=b, only=[0, 5])
    assert a[0].tobytes() == b[0].tobytes()
    # the pilot pass parks every replicate that reaches its depth (300 iterations here) and the full pass resumes
    # it: most of these replicates were interrupted once.  A pilot pass that is thrown away must give the same.
    assert int((a[0]["n_iterations"] > 300).sum()) > 24
    monkeypatch.delenv("JSB_NO_PILOT")
    monkeypatch.setenv("JSB_NO_SUSPEND", "1")
    c = []
    compare_runs(jsb, oracle, ("lattice", 2, 40, 40), p, seed=77, reps=48, warps_per_sm=2, summaries_out=c, only=[0, 5])
    assert a[0].tobytes() == c[0].tobytes()


def test_suspended_replicates_resume_in_tree_mode(jsb, oracle, monkeypatch):
    # method 2 (driver tree, hvoter displacements, the every-iteration excision index of :1062-1064) parked and resumed
    p = dict(method=2, Tf=60.0, rate_death=0.02, rate_migration=0.01, mu_driver=0.05, rule="hvoter", tree_edges=TREE_EDGES,
             node_rate=TREE_RATES, root_node=0, root_rate=0.3, t_bottleneck=[20.0, 40.0], ratio_bottleneck=[0.5, 0.5],
             t_snapshot=[15.0])
    monkeypatch.setenv("JSB_DEBUG_GRID", "3")
    monkeypatch.setenv("JSB_PILOT_ITERS", "256")
    got = []
    compare_runs(jsb, oracle, ("lattice", 3, 30, 30), p, seed=21, reps=36, warps_per_sm=2, summaries_out=got)
    assert int((got[0]["n_iterations"] > 256).sum()) > 15


def test_suspended_replicates_resume_with_snapshots_and_small_arena(jsb, oracle, monkeypatch):
    # resume with everything that lives across the interruption: snapshot index, excision index, event-log
    # chunks, a list arena that has been compacted, voter displacements
    p = dict(C2, Tf=60.0, rule="voter", t_bottleneck=[20.0, 40.0], ratio_bottleneck=[0.5, 0.5], t_snapshot=[5.0, 25.0, 55.0])
    monkeypatch.setenv("JSB_DEBUG_GRID", "3")
    monkeypatch.setenv("JSB_PILOT_ITERS", "256")
    monkeypatch.setenv("JSB_DEBUG_ARENA_ENTRIES", "16000")
    got = []
    compare_runs(jsb, oracle, ("lattice", 2, 36, 36), p, seed=5, reps=40, warps_per_sm=2, summaries_out=got)
    assert np.all(got[0]["status"] == 0) and int((got[0]["n_iterations"] > 256).sum()) > 15


@pytest.mark.parametrize("rule", ["contact", "hvoter"])
def test_c3_50cubed_tree_excision(jsb, oracle, rule):
    p = dict(method=2, Tf=200.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, rule=rule,
             tree_edges=TREE_EDGES, node_rate=TREE_RATES, root_node=0, root_rate=0.2,
             t_bottleneck=[100.0], ratio_bottleneck=[0.8])
    s = compare_runs(jsb, oracle, ("lattice", 3, 350, 350), p, seed=1234, reps=8, only=[0, 1, 2, 3])
    assert max(int(x["n_alive"]) for x in s) > 65535  # u32 site ids, lists longer than a u16 can index
    assert max(int(x["n_excised"]) for x in s) > 100 and max(int(x["n_clones"]) for x in s) == 4


def test_c4_100k_node_geometric_graph(jsb, oracle):
    from graphs import random_geometric_csr, geometric_centre
    rowptr, colidx, pts = random_geometric_csr(100000, seed=4321, return_points=True)
    centre = geometric_centre(pts)
    s = compare_runs(jsb, oracle, ("csr", rowptr, colidx), dict(C2, t_bottleneck=[], ratio_bottleneck=[]),
                     seed=1234, reps=8, candidates=[centre])
    assert len(rowptr) - 1 > 90000 and max(int(x["n_iterations"]) for x in s) > 10000


def test_csr_u32_replicates_parked_and_resumed(jsb, oracle, monkeypatch):
    # the large-state path (u32 site ids, CSR neighbour gathers, occupancy test in global memory) through the pilot:
    # more replicates than warp slots, every one parked once and resumed, eight of them against the oracle
    from graphs import random_geometric_csr, geometric_centre
    rowptr, colidx, pts = random_geometric_csr(70000, seed=99, return_points=True)
    assert len(rowptr) - 1 > 65535
    monkeypatch.setenv("JSB_DEBUG_GRID", "4")
    monkeypatch.setenv("JSB_PILOT_ITERS", "1000")
    got = []
    compare_runs(jsb, oracle, ("csr", rowptr, colidx), dict(C2, Tf=120.0, t_bottleneck=[60.0], ratio_bottleneck=[0.5]),
                 seed=4, reps=24, warps_per_sm=2, candidates=[geometric_centre(pts)], only=[0, 1, 2, 5, 8, 13, 21, 23],
                 summaries_out=got)
    assert int((got[0]["n_iterations"] > 1000).sum()) > 12


def test_c5_sweep_grid_corners(jsb, oracle):
    # b in {0.1..0.6} x d in {.005,.01,.02,.04} x adv_mean in {.1,.2,.4,.8}; the eight corners
    pts = [dict(Tf=200.0, rate_birth=b, rate_death=d, rate_migration=0.01, mu_driver=1e-5, adv_mean=a, adv_std=0.1,
                rule="contact") for b in (0.1, 0.6) for d in (0.005, 0.04) for a in (0.1, 0.8)]
    s = compare_runs(jsb, oracle, ("lattice", 2, 100, 100), pts, seed=99, reps=2)
    assert len(s) == 16 and max(int(x["n_iterations"]) for x in s) > 500_000
