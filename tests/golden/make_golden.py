"""Generates tests/golden/*.npz: outputs of the CPU oracle on small seeded cases.

The reference ships no golden vectors for this path and cannot run here (no Julia), so these
pin the ORACLE (regression) and give the GPU tests fixed files to compare with:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))

TREE = dict(method=2, tree_edges=[(0, 1), (1, 2), (2, 3)], node_rate=[0.2, 0.5, 0.55, 0.6], root_node=0, root_rate=0.2)
BASE = dict(Tf=60.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4, adv_std=0.1,
            rule="contact")
CASES = {
    # name: (graph spec, params, seed, stream)
    "contact_2d_32": (("lattice", 2, 32, 32), dict(BASE, Tf=95.0, t_bottleneck=[40.0], ratio_bottleneck=[0.8]), 1234, 0),
    "voter_2d_24": (("lattice", 2, 24, 24), dict(BASE, rule="voter", Tf=40.0, mu_driver=0.05), 1234, 1),
    "tree_3d_hvoter": (("lattice", 3, 30, 30), dict(BASE, **TREE, rule="hvoter", Tf=70.0, mu_driver=0.05,
                                                     t_bottleneck=[40.0], ratio_bottleneck=[0.8]), 1234, 2),
    "adjacency_10": (("dense", "Example_adj_matrix.txt"), dict(BASE, Tf=100.0, rate_death=0.05), 1234, 3),
}


def run_case(O, name):
    spec, params, seed, stream = CASES[name]
    if spec[0] == "lattice":
        g = O.Graph.lattice(spec[1], spec[2], spec[3])
    else:
        g = O.Graph.dense_file(os.path.join(HERE, spec[1]))
    r = O.Run(g, O.make_params(**params), seed, stream, faithful=True)
    return {"summary": np.array([r.summary]), "events": r.events, "clones": r.clones, "clone_of_site": r.clone_of_site,
            "cell_of_site": r.cell_of_site, "cs_alive": r.lists[0]}


if __name__ == "__main__":
    import oracle as O
    for name in CASES:
        out = run_case(O, name)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, dict(zip(out["summary"].dtype.names, out["summary"][0].tolist())))
