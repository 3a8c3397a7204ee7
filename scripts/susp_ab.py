"""A/B: C2 ensemble kernel time with the suspending pilot (default) and with a pilot pass that is thrown away."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J, numpy as np
from jspace_b200 import _lib as L
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()
ref = {}
for mode in ("suspend", "discard", "suspend32k", "suspend64k"):
    os.environ.pop("JSB_NO_SUSPEND", None); os.environ.pop("JSB_PILOT_ITERS", None)
    if mode == "discard": os.environ["JSB_NO_SUSPEND"] = "1"
    if mode == "suspend32k": os.environ["JSB_PILOT_ITERS"] = "32768"
    if mode == "suspend64k": os.environ["JSB_PILOT_ITERS"] = "65536"
    for flags, name in ((0, "summaries"), (L.OUT_EVENTS | L.OUT_LATTICE, "log+lattice")):
        ms = []
        for seed in (1234, 1235, 1236):
            r = J.run_ensemble(g, J.Point(**C2), 4096, seed=seed, flags=flags, log_pool_bytes=(14 << 30) if flags else 0)
            ms.append(r.kernel_ms)
            key = (seed, name)
            if key in ref: assert ref[key] == r.summaries.tobytes(), "summaries differ between pilot modes"
            else: ref[key] = r.summaries.tobytes()
            r.close()
        print(mode, name, "kernel ms", [round(x) for x in ms], flush=True)
