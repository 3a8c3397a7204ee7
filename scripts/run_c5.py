"""BASELINE config 5 at full size on one GPU: 128-point birth/death/advantage grid x 8192 replicates
= 2^20 replicates of the 100x100 lattice (Script1-style dynamics), summary-only."""
import sys, time
sys.path.insert(0, 'j-space.jl_b200')
import numpy as np
import jspace_b200 as J
from jspace_b200 import _lib as L
g = J.Graph.lattice(100, 100, 2)
pts = [J.Point(Tf=200.0, rate_birth=b, rate_death=d, rate_migration=0.01, mu_driver=1e-5, adv_mean=a, adv_std=0.1)
       for b in np.linspace(0.1, 0.6, 8) for d in (0.005, 0.01, 0.02, 0.04) for a in (0.1, 0.2, 0.4, 0.8)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for mode, flags in (("exact", 0), ("rejection-free", L.MODE_REJECTION_FREE)):
    t = time.time()
    r = J.run_ensemble(g, pts, reps, seed=2024, flags=flags)
    s = r.summaries
    print(mode, "replicates", r.count, "kernel_ms %.1f" % r.kernel_ms, "wall %.1f" % (time.time() - t),
          "iterations %.4g" % s["n_iterations"].sum(), "it/s %.3g" % (s["n_iterations"].sum() / (r.kernel_ms / 1e3)),
          "replicates/s %.0f" % (r.count / (r.kernel_ms / 1e3)), "status", np.bincount(s["status"]),
          "extinct %.3f" % (s["n_alive"] == 0).mean())
    r.close()
