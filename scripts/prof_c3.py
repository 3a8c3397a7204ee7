"""Debug: phase profile (JSB_PROFILE build, `make -C j-space.jl_b200 prof`) of single replicates of the C3 workload
(50^3 lattice, linear 4-driver tree, hvoter, excision); JSB_NO_HELPERS=1 for the unhelped critical path."""
import os, sys
os.environ['JSPACE_B200_LIB'] = os.path.abspath('j-space.jl_b200/lib/libjspace_b200_prof.so')
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
g = J.Graph.lattice(350, 350, 3)
p = J.Point(method=2, Tf=200.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, rule=sys.argv[1] if len(sys.argv) > 1 else "hvoter",
            tree_edges=[(0, 1), (1, 2), (2, 3)], node_rate=[0.2, 0.5, 0.55, 0.6], root_node=0, root_rate=0.2,
            t_bottleneck=[100.0], ratio_bottleneck=[0.8])
for s in (0, 1):
    r = J.run_ensemble(g, p, 256, seed=1234, g_begin=s, g_end=s + 1)
    sm = r.summaries[0]
    print("stream", s, "kernel_ms", round(r.kernel_ms), "it", int(sm['n_iterations']), "eff", int(sm['n_effective']), "births", int(sm['n_births']),
          "displaced", int(sm['n_displaced']), "deaths", int(sm['n_deaths']), "alive", int(sm['n_alive']), flush=True)
    r.close()
