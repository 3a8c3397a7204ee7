"""Debug: throughput of the exact kernel vs resident warps per SM (more warps = less L1 left by the smem carve-out)."""
import sys
sys.path.insert(0, 'j-space.jl_b200')
import jspace_b200 as J
g = J.Graph.lattice(200, 200, 2)
for wps in (10, 8, 6, 4, 2):
    reps = 148 * wps * 6
    r = J.run_ensemble(g, J.Point(Tf=150.0), reps, seed=1234, warps_per_sm=wps)
    it = r.summaries['n_iterations'].sum()
    print(f"warps/SM {wps:2d} reps {reps} kernel_ms {r.kernel_ms:8.1f} it {it:.3e} it/s {it / (r.kernel_ms / 1e3):.3e} per-warp {it / (r.kernel_ms / 1e3) / (148 * wps):.3e}", flush=True)
    r.close()
