"""Pins the oracle to the REFERENCE ITSELF (TEST INFRASTRUCTURE ONLY).

Runs the oracle's restatement of `Start_J_Space` -> `spatial_graph` -> `simulate_evolution` (method 1) on the
reference's own random stream (MersenneTwister(1234) restated in julia_mt.hpp) with the parameters the
reference ships (Parameters.toml / Config.toml: 25 x 25, dim = 3 -> 9^3 lattice, contact, Tf = 200, birth .2,
death .01, migration .01, driver rate .01, advantage N(.4, .1), excisions [50, 150] / [.8, .8], seed 1234) and
compares set_mut / alpha with the one output of this path the reference ships: fileoutput/DriverList.csv
(written by src/Start.jl:149-161).

    python oracle/pin_driverlist.py [path/to/DriverList.csv] [--search]

--search tries the alternatives of julia_mt.hpp's Variant knobs and reports which reproduce the file.

Status (round 2): the generator, the Float64 conversion, the ziggurat normal (fast path, wedge rejection) and the
fitness arithmetic ARE pinned against this file — tests/test_reference_pin.py finds all 73 fitness values bit for
bit in the stream of MersenneTwister(1234), in row order, each right after a driver-test draw <= 0.01.  The
whole-trajectory reproduction this script attempts is NOT achieved: the shipped file was not produced from the
shipped Parameters.toml (its sibling formatNewick holds 100 sampled cells, Parameters.toml asks for 10; the
neighbour draws before the 73 mutations use two random bits, i.e. a 2-D lattice, Parameters.toml says dim = 3),
and the lattice size / death / migration rates / excision schedule of that run cannot be read off the file."""
import ast
import csv
import ctypes as C
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402

SHIPPED = dict(method=1, Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4,
               adv_std=0.1, rule="contact", n_start_cells=1, t_bottleneck=[50.0, 150.0], ratio_bottleneck=[0.8, 0.8],
               t_snapshot=[10.0, 20.0])
DEFAULT_VARIANT = (0, 2, 0, 1, 1, 2, -1, 0)  # set_slots, sizehint_mode, hash_float, range_fast, set_ndl, int_fill_tail


def read_driverlist(path):
    rows = []
    with open(path, newline="") as f:
        for g, fit in csv.reader(f):
            geno = ast.literal_eval(g)
            rows.append(([geno] if isinstance(geno, int) else list(geno), float(fit)))
    return rows


def run_mt(seed=1234, variant=DEFAULT_VARIANT, row=25, col=25, dim=3, params=SHIPPED):
    L = O.lib()
    L.orc_run_mt.restype = C.c_void_p
    L.orc_run_mt.argtypes = [C.c_void_p, C.POINTER(O.OrcParams), C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64]
    L.orc_mt_randn_pos.restype = C.c_int64
    L.orc_mt_randn_pos.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    g = O.Graph.lattice(dim, row, col)
    p, keep = O.make_params(**params)
    v = np.asarray(variant, dtype=np.int32)
    h = L.orc_run_mt(g.h, C.byref(p), seed, v.ctypes.data, None, 0)
    try:
        ptr = C.c_void_p()
        n = L.orc_clones(h, C.byref(ptr))
        clones = np.frombuffer(C.string_at(ptr.value, n * 16), dtype=O.CLONE_DTYPE).copy()
        summary = np.frombuffer(C.string_at(L.orc_summary(h), 64), dtype=O.SUMMARY_DTYPE)[0].copy()
        n = L.orc_mt_randn_pos(h, C.byref(ptr))
        run_mt.randn_pos = np.frombuffer(C.string_at(ptr.value, n * 8), dtype=np.int64).copy() if n else np.zeros(0, np.int64)
    finally:
        L.orc_free(h)
    genos = []
    for c in clones:
        genos.append((genos[c["parent"] - 1] if c["parent"] else []) + [int(c["label"])])
    return [(g, float(c["alpha"])) for g, c in zip(genos, clones)], summary


def compare(ours, ref):
    """(rows equal in genotype, rows equal in genotype and fitness bit for bit, first mismatch)"""
    n_g = n_f = 0
    first = None
    for i, (a, b) in enumerate(zip(ours, ref)):
        if a[0] == b[0]:
            n_g += 1
            if a[1] == b[1]:
                n_f += 1
            elif first is None:
                first = (i, a, b)
        elif first is None:
            first = (i, a, b)
    return n_g, n_f, first


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    path = args[0] if args else os.path.join(HERE, "..", "tests", "golden", "reference_DriverList.csv")
    ref = read_driverlist(path)
    O.build()
    if "--search" in sys.argv:
        for var in itertools.product((0, 1024, 2048, 4096), (0, 1, 2), (0, 1), (0, 1), (0, 1), (0, 2), (-1,), (0,)):
            ours, s = run_mt(variant=var)
            n_g, n_f, first = compare(ours, ref)
            if n_g > 1:
                print(var, "clones", len(ours), "genotypes equal", n_g, "fitness equal", n_f, "first mismatch", first)
        return 0
    ours, s = run_mt()
    n_g, n_f, first = compare(ours, ref)
    print(f"oracle on MersenneTwister(1234): {len(ours)} clones, {int(s['n_iterations'])} iterations; reference file: {len(ref)} rows")
    print(f"rows with the same genotype: {n_g}; same genotype and bit-identical fitness: {n_f}; first mismatch: {first}")
    for row in ours[:6]:
        print("  ", row)
    print("raw dSFMT positions of our randn() calls:", run_mt.randn_pos[:12].tolist())
    return 0 if (len(ours) == len(ref) and n_f == len(ref)) else 1


if __name__ == "__main__":
    sys.exit(main())
