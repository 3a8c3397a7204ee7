import sys, time
sys.path.insert(0,'j-space.jl_b200')
import jspace_b200 as J, numpy as np
from jspace_b200 import _lib as L
g=J.Graph.lattice(200,200,2)
pt=J.Point(t_bottleneck=[50.0], ratio_bottleneck=[0.8])
for reps in (256, 4096):
    r=J.run_ensemble(g, pt, reps, seed=1234, flags=L.MODE_REJECTION_FREE)
    s=r.summaries
    print("fast", reps, "kernel_ms", round(r.kernel_ms,1), "events", int(s['n_iterations'].sum()), "ev/s %.3g"%(s['n_iterations'].sum()/(r.kernel_ms/1e3)), "reps/s %.1f"%(reps/(r.kernel_ms/1e3)), "alive mean", s['n_alive'].mean(), "clones mean", s['n_clones'].mean(), "status", np.bincount(s['status']))
    r.close()
