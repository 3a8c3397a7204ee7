"""Debug: the same 148 replicates, one per SM vs packed ten per SM (JSB_DEBUG_GRID)."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
g = J.Graph.lattice(200, 200, 2)
r = J.run_ensemble(g, J.Point(), 148, seed=1234)
it = r.summaries['n_iterations']
print("grid", os.environ.get("JSB_DEBUG_GRID", "default"), "kernel_ms", r.kernel_ms, "total it", it.sum(), "max it", it.max(), "it/s", it.sum() / (r.kernel_ms / 1e3))
