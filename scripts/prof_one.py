"""Debug: phase profile (JSB_PROFILE build) of single heavy replicates of the C2 ensemble."""
import os, sys
os.environ['JSPACE_B200_LIB'] = os.path.abspath('j-space.jl_b200/lib/libjspace_b200_prof.so')
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
g = J.Graph.lattice(200, 200, 2)
for s in [int(a) for a in sys.argv[1:]] or [2956]:
    r = J.run_ensemble(g, J.Point(), 4096, seed=1234, g_begin=s, g_end=s + 1)
    print("stream", s, "kernel_ms", r.kernel_ms, "it", r.summaries['n_iterations'].sum(), flush=True)
    r.close()
