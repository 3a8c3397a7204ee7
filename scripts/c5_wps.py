"""Debug: a reduced BASELINE config 5 sweep (128 points x 512 replicates, 100x100) vs resident warps per SM."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import numpy as np
import jspace_b200 as J
g = J.Graph.lattice(100, 100, 2)
pts = [J.Point(Tf=200.0, rate_birth=b, rate_death=d, rate_migration=0.01, mu_driver=1e-5, adv_mean=a, adv_std=0.1)
       for b in np.linspace(0.1, 0.6, 8) for d in (0.005, 0.01, 0.02, 0.04) for a in (0.1, 0.2, 0.4, 0.8)]
for wps in [int(a) for a in sys.argv[1:]] or [12, 10, 8]:
    r = J.run_ensemble(g, pts, 512, seed=2024, warps_per_sm=wps)
    it = r.summaries["n_iterations"].sum()
    print("warps/SM", wps, "kernel_ms %.1f" % r.kernel_ms, "it/s %.3e" % (it / (r.kernel_ms / 1e3)), flush=True)
    r.close()
