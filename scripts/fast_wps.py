"""Debug: the rejection-free kernel on the C2 ensemble vs resident warps per SM."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
from jspace_b200 import _lib as L
g = J.Graph.lattice(200, 200, 2)
pt = J.Point(t_bottleneck=[50.0], ratio_bottleneck=[0.8])
for wps in [int(a) for a in sys.argv[1:]] or [12, 10, 8, 7, 6]:
    r = J.run_ensemble(g, pt, 4096, seed=1234, flags=L.MODE_REJECTION_FREE, warps_per_sm=wps)
    print("warps/SM", wps, "kernel_ms", round(r.kernel_ms, 1), "replicates/s %.0f" % (4096 / (r.kernel_ms / 1e3)), flush=True)
    r.close()
