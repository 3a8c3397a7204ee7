import sys
sys.path.insert(0,'j-space.jl_b200')
import jspace_b200 as J
reps=int(sys.argv[1]) if len(sys.argv)>1 else 148
Tf=float(sys.argv[2]) if len(sys.argv)>2 else 14.0
g=J.Graph.lattice(100,100,2)
# fast growth fills the lattice by t~8; afterwards ~99% null iterations with a few hundred clones
r=J.run_ensemble(g, J.Point(Tf=Tf, rate_birth=3.0, rate_death=0.02, rate_migration=0.02, mu_driver=0.04, adv_mean=0.05, adv_std=0.02), reps, seed=1)
s=r.summaries
print(reps, "kernel_ms", r.kernel_ms, "iters", s['n_iterations'].sum(), "eff", s['n_effective'].sum(), "clones", s['n_clones'].mean(), "alive", s['n_alive'].mean())
