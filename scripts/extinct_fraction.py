import sys
sys.path.insert(0, 'j-space.jl_b200'); sys.path.insert(0,'.')
import jspace_b200 as J, numpy as np, bench
C2 = bench.WORKLOADS["C2"]["points"][0]
g = J.Graph.lattice(200, 200, 2)
r = J.run_ensemble(g, J.Point(**C2), 4096, seed=1234)
sm = r.summaries
heavy = np.argsort(-sm["n_iterations"].astype(np.int64))[:40]
fr = []
for i in heavy:
    c = r.clones(int(i))
    fr.append(((c["count"] == 0).mean(), len(c), int(np.flatnonzero(c["count"] > 0).min()) if (c["count"] > 0).any() else -1))
print("heaviest 40: extinct fraction mean %.3f min %.3f max %.3f; clones mean %.0f; first live index median %d" % (
    np.mean([f[0] for f in fr]), np.min([f[0] for f in fr]), np.max([f[0] for f in fr]), np.mean([f[1] for f in fr]), np.median([f[2] for f in fr])))
allf = [(r.clones(i)["count"] == 0).mean() for i in range(0, 4096, 16) if sm["n_clones"][i] > 10]
print("all (sample): extinct fraction mean %.3f" % np.mean(allf))
