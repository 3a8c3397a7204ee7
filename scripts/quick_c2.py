import sys, time
sys.path.insert(0,'j-space.jl_b200')
import jspace_b200 as J, numpy as np
g=J.Graph.lattice(200,200,2)
for reps in (256, 4096):
    t=time.time()
    r=J.run_ensemble(g, J.Point(), reps, seed=1234)
    wall=time.time()-t
    it=r.summaries['n_iterations'].sum()
    print(reps, "kernel_ms", r.kernel_ms, "wall", wall, "iters", it, "it/s", it/(r.kernel_ms/1e3), "clones mean", r.summaries['n_clones'].mean(), "alive mean", r.summaries['n_alive'].mean(), "status", np.bincount(r.summaries['status']))
    r.close()
