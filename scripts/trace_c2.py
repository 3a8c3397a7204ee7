"""Debug: per-replicate timeline of the exact kernel on the C2 ensemble (JSB_TRACE hook).
Writes gpurun_out/trace_c2.bin (u64[4] per replicate: start ns, end ns, SM<<8|warp, iterations)."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
path = 'gpurun_out/trace_c2.bin'
os.makedirs('gpurun_out', exist_ok=True)
os.environ['JSB_TRACE'] = path
import jspace_b200 as J, numpy as np
C2 = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
          rule="contact", t_bottleneck=[50.0], ratio_bottleneck=[0.8])
g = J.Graph.lattice(200, 200, 2)
J.run_ensemble(g, J.Point(**C2), 256, seed=1).close()
r = J.run_ensemble(g, J.Point(**C2), 4096, seed=int(os.environ.get("SEED", "1234")))
print("kernel_ms", r.kernel_ms)
np.save('gpurun_out/trace_c2_summaries.npy', r.summaries)
tr = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)[-4096:]
t0 = tr[:, 0].min()
st = (tr[:, 0] - t0) / 1e6
en = (tr[:, 1] - t0) / 1e6
it = tr[:, 3].astype(np.float64)
sm = tr[:, 2] >> 8
print("makespan ms", en.max(), "last start ms", st.max())
order = np.argsort(-it)
for i in order[:12]:
    print(f"rep {i} it {it[i]:.3e} start {st[i]:.0f} end {en[i]:.0f} dur {en[i]-st[i]:.0f} sm {sm[i]} rate {it[i]/(en[i]-st[i])/1e3:.2f} M it/s")
grid = np.arange(0, en.max(), 100.0)
busy = [int(np.sum((st <= t) & (en > t))) for t in grid]
print("busy warps every 100 ms:", busy)
late = np.argsort(-en)[:16]
print("last finishers (rep, iterations, start, end):", [(int(i), f"{it[i]:.2e}", f"{st[i]:.0f}", f"{en[i]:.0f}") for i in late])
done = [float(np.sum(it * np.clip((t - st) / np.maximum(en - st, 1e-9), 0, 1))) for t in grid]
print("cumulative iterations (1e9) every 100 ms:", [round(d / 1e9, 2) for d in done])
