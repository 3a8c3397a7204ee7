"""Debug: C2 ensemble at different numbers of resident warps per SM (same box)."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
g = J.Graph.lattice(200, 200, 2)
for wps in [int(a) for a in sys.argv[1:]] or [10, 12, 10, 12, 11]:
    r = J.run_ensemble(g, J.Point(), 4096, seed=1234, warps_per_sm=wps)
    it = r.summaries['n_iterations'].sum()
    print("warps/SM", wps, "kernel_ms", round(r.kernel_ms, 1), "it/s %.3e" % (it / (r.kernel_ms / 1e3)), flush=True)
    r.close()
