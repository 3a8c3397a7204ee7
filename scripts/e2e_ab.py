"""Debug: kernel time of the C2 ensemble with the event log and the lattices requested."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
from jspace_b200 import _lib as L
g = J.Graph.lattice(200, 200, 2)
pt = J.Point(t_bottleneck=[50.0], ratio_bottleneck=[0.8])
for s in (1234, 1235):
    r = J.run_ensemble(g, pt, 4096, seed=s, flags=L.OUT_EVENTS | L.OUT_LATTICE, log_pool_bytes=14 << 30)
    print("seed", s, "kernel_ms", round(r.kernel_ms, 1), flush=True)
    r.close()
