"""A/B: one heavy C2 replicate alone on the device with 0, 1, 2, ... helper warps (clock helper first, or lookahead only)."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J, numpy as np
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()
for noclock in (0, 1):
    if noclock: os.environ["JSB_NO_CLOCK"] = "1"
    for s in (3774, 2956, 1513):
        row = []
        for wps in (1, 2, 3, 4, 6):
            r = J.run_ensemble(g, J.Point(**C2), 4096, seed=1234, g_begin=s, g_end=s + 1, warps_per_sm=wps)
            row.append(round(r.kernel_ms)); r.close()
        print("no_clock" if noclock else "clock   ", "stream", s, "warps 1,2,3,4,6 ->", row, flush=True)
