"""Shared helper: run the same replicates through the CUDA library (C ABI) and the CPU oracle
and compare every output bit for bit.  Test infrastructure only."""
from __future__ import annotations

import numpy as np

TREE_EDGES = [(0, 1), (1, 2), (2, 3)]          # utility/Tree_driver_example.txt (linear 4-driver tree)
TREE_RATES = [0.2, 0.5, 0.55, 0.6]             # utility/driver_advantage.txt


def make_graphs(jsb, oracle, spec):
    kind = spec[0]
    if kind == "lattice":
        _, dim, row, col = spec
        return jsb.Graph.lattice(row, col, dim), oracle.Graph.lattice(dim, row, col)
    if kind == "csr":
        _, rowptr, colidx = spec
        return jsb.Graph.csr(rowptr, colidx), oracle.Graph.csr(rowptr, colidx)
    if kind == "dense":
        return jsb.Graph.dense_file(spec[1]), oracle.Graph.dense_file(spec[1])
    raise ValueError(kind)


def compare_runs(jsb, oracle, spec, params: dict, *, seed=1234, reps=1, g_begin=0, faithful=False,
                 time_rtol=0.0, candidates=None, check_lists=True, only=None, warps_per_sm=0, log_pool_bytes=0,
                 summaries_out=None):
    """Returns the list of oracle summaries; raises AssertionError on the first mismatch.
    `only`: local replicate indices to compare with the oracle (default: all of them) — the GPU
    still runs the whole ensemble, so scheduling (pilot order, helpers, decoupled clock) is live."""
    from jspace_b200 import _lib as L
    gj, go = make_graphs(jsb, oracle, spec)
    if candidates is not None:
        gj.set_seed_candidates(candidates)
    flags = L.OUT_EVENTS | L.OUT_LATTICE | (L.OUT_LISTS if check_lists else 0)
    # `params` may be a list of dicts = sweep points; then `reps` is the number of replicates PER POINT and
    # replicate g = point * reps + r runs point g // reps (the library's global numbering)
    plist = params if isinstance(params, (list, tuple)) else None
    reps_per_point = reps if plist else g_begin + reps
    if plist:
        assert g_begin == 0
        reps = reps * len(plist)
    point = [jsb.Point(**q) for q in plist] if plist else jsb.Point(**params)
    res = jsb.run_ensemble(gj, point, reps_per_point, seed=seed, flags=flags, g_begin=g_begin,
                           g_end=g_begin + reps, warps_per_sm=warps_per_sm, log_pool_bytes=log_pool_bytes)
    if summaries_out is not None:
        summaries_out.append(res.summaries.copy())
    cand = candidates
    if cand is None and spec[0] != "lattice":
        cand = gj.seed_candidates()
    out = []
    for i in (range(res.count) if only is None else only):
        stream = g_begin + i
        params = plist[stream // reps_per_point] if plist else params
        orun = oracle.Run(go, oracle.make_params(**params), seed, stream, candidates=cand, faithful=faithful)
        s_gpu, s_cpu = res.summaries[i], orun.summary
        ctx = f"replicate {stream} params {params}"
        for f in s_cpu.dtype.names:
            if f == "t_final" and time_rtol:
                assert np.isclose(s_gpu[f], s_cpu[f], rtol=time_rtol, atol=0), (ctx, f, s_gpu[f], s_cpu[f])
            else:
                assert s_gpu[f] == s_cpu[f], (ctx, f, s_gpu[f], s_cpu[f])
        ev_gpu, ev_cpu = res.events(i), orun.events
        assert len(ev_gpu) == len(ev_cpu), (ctx, len(ev_gpu), len(ev_cpu))
        for f in ("cell", "parent", "site", "site2", "clone", "label", "type", "flags", "pad"):
            bad = np.nonzero(ev_gpu[f] != ev_cpu[f])[0]
            assert bad.size == 0, (ctx, f, int(bad[0]), ev_gpu[bad[0]], ev_cpu[bad[0]])
        if time_rtol:
            assert np.allclose(ev_gpu["time"], ev_cpu["time"], rtol=time_rtol, atol=0), ctx
        else:
            bad = np.nonzero(ev_gpu["time"] != ev_cpu["time"])[0]
            assert bad.size == 0, (ctx, "time", int(bad[0]), ev_gpu[bad[0]], ev_cpu[bad[0]])
        cl_gpu, cl_cpu = res.clones(i), orun.clones
        assert len(cl_gpu) == len(cl_cpu), ctx
        for f in ("count", "parent", "label"):
            assert np.array_equal(cl_gpu[f], cl_cpu[f]), (ctx, f)
        if time_rtol:
            assert np.allclose(cl_gpu["alpha"], cl_cpu["alpha"], rtol=time_rtol, atol=0), ctx
        else:
            assert np.array_equal(cl_gpu["alpha"], cl_cpu["alpha"]), (ctx, "alpha", cl_gpu["alpha"], cl_cpu["alpha"])
        lat_c, lat_i = res.lattice(i)
        assert np.array_equal(lat_c, orun.clone_of_site), (ctx, "lattice clone")
        assert np.array_equal(lat_i, orun.cell_of_site), (ctx, "lattice cell ids")
        if check_lists:
            for li in range(int(s_cpu["n_clones"]) + 1):
                assert np.array_equal(res.list(i, li), orun.lists[li]), (ctx, "list", li)
        sn_gpu, sn_cpu = res.snapshots(i), orun.snapshots
        assert len(sn_gpu) == len(sn_cpu), (ctx, "snapshots")
        for f in ("iteration", "n_events_before", "n_alive", "index"):
            assert np.array_equal(sn_gpu[f], sn_cpu[f]), (ctx, "snapshot", f)
        if not time_rtol:
            assert np.array_equal(sn_gpu["time"], sn_cpu["time"]), (ctx, "snapshot time")
        out.append(s_cpu)
    res.close()
    return out
