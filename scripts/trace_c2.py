"""Debug: per-replicate timeline of the exact kernel on the C2 ensemble (JSB_TRACE hook)."""
import os, sys
sys.path.insert(0, 'j-space.jl_b200')
path = 'gpurun_out/trace_c2.bin'
os.makedirs('gpurun_out', exist_ok=True)
os.environ['JSB_TRACE'] = path
import jspace_b200 as J, numpy as np
g = J.Graph.lattice(200, 200, 2)
r = J.run_ensemble(g, J.Point(), 4096, seed=1234)
print("kernel_ms", r.kernel_ms)
tr = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)
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
busy = [(np.sum((st <= t) & (en > t))) for t in grid]
print("busy warps every 100 ms:", busy)
late = np.argsort(-en)[:12]
print("last finishers:", [(int(i), f"{it[i]:.2e}", f"{st[i]:.0f}", f"{en[i]:.0f}") for i in late])
# work done over time assuming uniform rate within a replicate
done = [float(np.sum(it * np.clip((t - st) / np.maximum(en - st, 1e-9), 0, 1))) for t in grid]
print("cumulative iterations (1e9) every 100 ms:", [round(d / 1e9, 2) for d in done])
