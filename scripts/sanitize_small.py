"""compute-sanitizer target: a small ensemble through every round-2 path (suspending pilot, resume, streamed outputs on the
second call, series reduction, log-pool exhaustion).  usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
os.environ["JSB_DEBUG_GRID"] = "4"
os.environ["JSB_PILOT_ITERS"] = "300"
import jspace_b200 as J
from jspace_b200 import _lib as L
g = J.Graph.lattice(40, 40, 2)
p = J.Point(Tf=70.0, rate_birth=0.3, rate_death=0.02, rate_migration=0.01, mu_driver=0.02, adv_mean=0.4, adv_std=0.1, rule="voter",
            t_bottleneck=[30.0], ratio_bottleneck=[0.5], t_snapshot=[10.0, 40.0])
flags = L.OUT_EVENTS | L.OUT_LATTICE | L.OUT_LISTS | L.OUT_SERIES
tot = 0
for call in range(3):
    r = J.run_ensemble(g, p, 48, seed=5 + call, flags=flags, warps_per_sm=2, log_pool_bytes=(8 << 20) if call < 2 else 40 * 32768)
    tot += int(r.summaries["n_events"].sum()) + len(r.events(3)) + len(r.series_rows(3)[0])
    print("call", call, "launches", r.launches, "statuses", sorted(set(r.summaries["status"].tolist())), flush=True)
    r.close()
J._lib.lib().jsb_trim()
print("sanitize_small ok", tot)
