"""A/B: kernel time of the C2 ensemble against the pilot depth (JSB_PILOT_ITERS) and warps per SM."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J, numpy as np
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()
for D in (8192, 16384, 32768, 65536):
    os.environ["JSB_PILOT_ITERS"] = str(D)
    ms = []
    for seed in (1234, 1235, 1236):
        r = J.run_ensemble(g, J.Point(**C2), 4096, seed=seed)
        ms.append(r.kernel_ms); r.close()
    print("pilot", D, "kernel ms", [round(x) for x in ms], flush=True)
os.environ["JSB_PILOT_ITERS"] = "32768"
for wps in (6, 8):
    ms = []
    for seed in (1234, 1235):
        r = J.run_ensemble(g, J.Point(**C2), 4096, seed=seed, warps_per_sm=wps)
        ms.append(r.kernel_ms); r.close()
    print("pilot 32768 wps", wps, "kernel ms", [round(x) for x in ms], flush=True)
