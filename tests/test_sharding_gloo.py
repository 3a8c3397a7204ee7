"""world_size-2 gloo test of the N>1 path's host logic: each rank takes its shard of the replicate
range, the per-rank summary histograms are summed with a collective, and the result equals the
single-process histogram (replicate g uses stream g wherever it runs).  The per-replicate
summaries come from the CPU oracle here (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL = 24
PARAMS = dict(Tf=40.0, rate_birth=0.25, rate_death=0.02, rate_migration=0.01, mu_driver=0.02, adv_mean=0.4,
              adv_std=0.1, rule="contact")


def _summaries(g_begin, g_end):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    g = O.Graph.lattice(2, 20, 20)
    return np.array([O.Run(g, O.make_params(**PARAMS), 77, s, faithful=False, want=()).summary
                     for s in range(g_begin, g_end)])


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "j-space.jl_b200"))
    from jspace_b200.distributed import allreduce_histograms, histograms_from_summaries, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    a, b = shard_range(rank, world, TOTAL)
    h = allreduce_histograms(histograms_from_summaries(_summaries(a, b), 400))
    if rank == 0:
        np.save(out, h)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_histogram_gather_equals_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "j-space.jl_b200"))
    from jspace_b200.distributed import histograms_from_summaries
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "h.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    want = histograms_from_summaries(_summaries(0, TOTAL), 400)
    assert np.array_equal(got, want) and got[0].sum() == TOTAL
