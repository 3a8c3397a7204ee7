"""Synthetic graphs for tests and benchmarks (BASELINE config 4: random geometric graph)."""
import numpy as np


def random_geometric_csr(n, seed=4321, mean_degree=8.0, return_points=False):
    """n points uniform in the unit square, edge when distance < r, r = sqrt(mean_degree/(pi n)).
    Returns (rowptr, colidx) of the largest connected component, 1-based ascending neighbours
    (and, with return_points, the coordinates of its vertices)."""
    rng = np.random.Generator(np.random.Philox(seed))
    pts = rng.random((n, 2))
    r = np.sqrt(mean_degree / (np.pi * n))
    cell = np.floor(pts / r).astype(np.int64)
    ncell = int(np.ceil(1.0 / r)) + 1
    key = cell[:, 0] * ncell + cell[:, 1]
    order = np.argsort(key, kind="stable")
    skey = key[order]
    src, dst = [], []
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            nk = (cell[:, 0] + dx) * ncell + (cell[:, 1] + dy)
            lo = np.searchsorted(skey, nk, "left")
            hi = np.searchsorted(skey, nk, "right")
            cnt = hi - lo
            idx = np.repeat(np.arange(n), cnt)
            off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            other = order[np.repeat(lo, cnt) + off]
            d2 = ((pts[idx] - pts[other]) ** 2).sum(1)
            m = (d2 < r * r) & (idx != other)
            src.append(idx[m]); dst.append(other[m])
    src = np.concatenate(src); dst = np.concatenate(dst)
    # largest component by label propagation on the undirected edge list
    label = np.arange(n)
    while True:
        new = label.copy()
        np.minimum.at(new, src, label[dst])
        new = np.minimum(new, new[new])
        if np.array_equal(new, label):
            break
        label = new
    vals, counts = np.unique(label, return_counts=True)
    keep = label == vals[np.argmax(counts)]
    remap = -np.ones(n, dtype=np.int64)
    remap[keep] = np.arange(keep.sum())
    m = keep[src] & keep[dst]
    s, d = remap[src[m]], remap[dst[m]]
    o = np.lexsort((d, s))
    s, d = s[o], d[o]
    nv = int(keep.sum())
    rowptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(rowptr, s + 1, 1)
    rowptr = np.cumsum(rowptr)
    if return_points:
        return rowptr, (d + 1).astype(np.int32), pts[keep]
    return rowptr, (d + 1).astype(np.int32)


def geometric_centre(pts):
    """1-based vertex nearest to the middle of the unit square: the stand-in for the max-closeness vertex of
    src/J_Space.jl:83-85 on graphs where the exact O(V*E) closeness is too slow to compute in a test."""
    return int(np.argmin(((pts - 0.5) ** 2).sum(1))) + 1
