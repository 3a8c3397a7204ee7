"""Debug: how well does a work estimate taken after D iterations predict a replicate's cost?
Runs the C2 ensemble capped at D iterations for several D and stores the estimate
lambda * (Tf - t) of every replicate, next to the JSB_TRACE timeline of the full run."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
os.makedirs('gpurun_out', exist_ok=True)
os.environ['JSB_TRACE'] = 'gpurun_out/trace_c2.bin'
import jspace_b200 as J, numpy as np
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
seed = int(os.environ.get("SEED", "1234"))
r = J.run_ensemble(g, J.Point(**C2), 4096, seed=seed)
print("full kernel_ms", r.kernel_ms)
out = {"full_it": r.summaries["n_iterations"].copy(), "full_eff": r.summaries["n_effective"].copy()}
r.close()
tr = np.fromfile('gpurun_out/trace_c2.bin', dtype=np.uint64).reshape(-1, 4)[-4096:]
out["trace"] = tr
del os.environ['JSB_TRACE']
for D in (8192, 32768, 131072, 524288, 2097152):
    r = J.run_ensemble(g, J.Point(**dict(C2, max_iterations=D)), 4096, seed=seed)
    sm = r.summaries
    prio = np.zeros(4096)
    for i in range(4096):
        if sm["n_alive"][i] == 0 or sm["t_final"][i] >= 200.0:
            continue
        c = r.clones(i)
        lam = float((c["alpha"] * c["count"]).sum()) + 0.02 * float(sm["n_alive"][i])
        prio[i] = lam * (200.0 - float(sm["t_final"][i]))
    out[f"prio_{D}"] = prio
    out[f"t_{D}"] = sm["t_final"].copy()
    out[f"it_{D}"] = sm["n_iterations"].copy()
    print(D, "kernel_ms", r.kernel_ms, "iterations", int(sm["n_iterations"].sum()))
    r.close()
np.savez('gpurun_out/pilot_quality.npz', **out)
