import sys, time
sys.path.insert(0,'j-space.jl_b200')
import jspace_b200 as J, numpy as np
reps=int(sys.argv[1]) if len(sys.argv)>1 else 592
Tf=float(sys.argv[2]) if len(sys.argv)>2 else 120.0
g=J.Graph.lattice(200,200,2)
r=J.run_ensemble(g, J.Point(Tf=Tf), reps, seed=1234)
it=r.summaries['n_iterations'].sum()
print(reps, Tf, "kernel_ms", r.kernel_ms, "iters", it, "it/s", it/(r.kernel_ms/1e3), "eff", r.summaries['n_effective'].sum(), "clones mean", r.summaries['n_clones'].mean(), "alive mean", r.summaries['n_alive'].mean())
