import sys, time, os
sys.path.insert(0,'j-space.jl_b200')
import jspace_b200 as J
from jspace_b200 import _lib as L
g=J.Graph.lattice(200,200,2)
pt=J.Point(t_bottleneck=[50.0], ratio_bottleneck=[0.8])
for s in (2000, 1234, 1235):
    t=time.time()
    r=J.run_ensemble(g, pt, 4096, seed=s, flags=L.OUT_EVENTS|L.OUT_LATTICE, log_pool_bytes=14<<30)
    t1=time.time()
    it=int(r.summaries['n_iterations'].sum()); ev=int(r.summaries['n_events'].sum())
    r.close()
    print("seed",s,"wall",round(t1-t,3),"close",round(time.time()-t1,3),"kernel_ms",round(r.kernel_ms,1),"events",ev, "GB", ev*32/1e9, file=sys.stderr)
