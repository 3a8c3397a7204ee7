"""Sweep of the tail knobs (environment hooks) on the C2 ensemble: clock helper admission, decoupled-mode threshold."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()
def run(tag, env, wps=0):
    for k in ("JSB_CLOCK_MAX_BUSY", "JSB_DEC_MIN_CLONES"): os.environ.pop(k, None)
    os.environ.update(env)
    ms = []
    for seed in (1234, 1235, 1236):
        r = J.run_ensemble(g, J.Point(**C2), 4096, seed=seed, warps_per_sm=wps)
        ms.append(r.kernel_ms); r.close()
    print(tag, [round(x) for x in ms], "mean", round(sum(ms) / len(ms)), flush=True)
run("default (busy<=3, clones>=64)", {})
for b in (1, 2, 4, 5, 7):
    run(f"clock_max_busy={b}", {"JSB_CLOCK_MAX_BUSY": str(b)})
for c in (16, 32, 128, 256):
    run(f"dec_min_clones={c}", {"JSB_DEC_MIN_CLONES": str(c)})
run("8 warps", {}, wps=8)
run("default again", {})
