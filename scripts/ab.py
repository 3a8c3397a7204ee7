"""A/B measurements of a library build (JSPACE_B200_LIB selects it): kernel times of
  full  the C2 ensemble (4096 x 200x200), summaries only and with log+lattice
  thr   the pure throughput regime (every warp slot busy, replicates capped at 600 k iterations)
  one   the heaviest C2 replicates alone (helpers + decoupled clock)
  c3    64 replicates of the C3 workload (50^3, tree, hvoter, excision)
usage: python scripts/ab.py full thr one c3"""
import sys, time
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
from jspace_b200 import _lib as L
import numpy as np
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
what = sys.argv[1:] or ["full", "thr", "one", "c3"]
tag = L.LIB_PATH.split('/')[-1]
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()  # warm-up
if "full" in what:
    for flags, name in ((0, "summaries"), (L.OUT_EVENTS | L.OUT_LATTICE, "log+lattice")):
        for rep in range(2):
            t0 = time.perf_counter()
            r = J.run_ensemble(g, J.Point(**C2), 4096, seed=1234 + rep, flags=flags, log_pool_bytes=(14 << 30) if flags else 0)
            wall = time.perf_counter() - t0
            it = int(r.summaries['n_iterations'].sum())
            print(f"[{tag}] full {name} seed {1234+rep}: kernel {r.kernel_ms:.1f} ms wall {wall*1e3:.0f} ms it {it:.3e} -> {it/(r.kernel_ms/1e3):.3e} it/s", flush=True)
            r.close()
if "thr" in what:
    for wps in (7,):
        r = J.run_ensemble(g, J.Point(**dict(C2, max_iterations=600000)), 148 * wps, seed=1234, warps_per_sm=wps)
        it = r.summaries['n_iterations'].sum()
        print(f"[{tag}] thr wps {wps}: kernel {r.kernel_ms:.1f} ms it {it:.3e} -> {it/(r.kernel_ms/1e3):.3e} it/s", flush=True)
        r.close()
if "one" in what:
    for s in (13, 6, 2956):
        r = J.run_ensemble(g, J.Point(**C2), 4096, seed=1234, g_begin=s, g_end=s + 1)
        sm = r.summaries[0]
        print(f"[{tag}] one stream {s}: kernel {r.kernel_ms:.1f} ms it {int(sm['n_iterations']):.3e} clones {sm['n_clones']} eff {sm['n_effective']} deaths {sm['n_deaths']} migr {sm['n_migrations']}", flush=True)
        r.close()
if "c3" in what:
    g3 = J.Graph.lattice(350, 350, 3)
    p = J.Point(method=2, Tf=200.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, rule="hvoter",
                tree_edges=[(0, 1), (1, 2), (2, 3)], node_rate=[0.2, 0.5, 0.55, 0.6], root_node=0, root_rate=0.2,
                t_bottleneck=[100.0], ratio_bottleneck=[0.8])
    r = J.run_ensemble(g3, p, 64, seed=1234)
    it = r.summaries['n_iterations'].sum()
    print(f"[{tag}] c3 64 reps: kernel {r.kernel_ms:.1f} ms it {it:.3e} -> {it/(r.kernel_ms/1e3):.3e} it/s eff {r.summaries['n_effective'].sum():.3e}", flush=True)
    r.close()
