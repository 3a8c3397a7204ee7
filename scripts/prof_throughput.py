"""Debug/profiling: the exact kernel in its pure throughput regime — every warp slot busy with the same
amount of work (replicates capped at the same iteration count)."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
wps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 600000
g = J.Graph.lattice(200, 200, 2)
r = J.run_ensemble(g, J.Point(max_iterations=cap), 148 * wps, seed=1234, warps_per_sm=wps)
it = r.summaries['n_iterations'].sum()
print("warps/SM", wps, "cap", cap, "kernel_ms", round(r.kernel_ms, 1), "it %.3e" % it, "it/s %.3e" % (it / (r.kernel_ms / 1e3)),
      "capped", int((r.summaries['status'] == 4).sum()), "of", 148 * wps, flush=True)
