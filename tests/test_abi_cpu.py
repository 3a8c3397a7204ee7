"""The C-ABI library without a GPU: it loads, exports every symbol include/jspace_b200.h declares,
validates arguments, builds graphs exactly like the oracle's restatement of Graphs.jl, and fails
loudly (JSB_ENODEVICE) instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "jspace_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(jsb_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(jsb):
    from jspace_b200 import _lib as L
    lib = L.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(L.SIGNATURES) == syms  # the ctypes binding covers exactly the header
    assert lib.jsb_abi_version() == 2


def test_struct_layouts_match_the_header(jsb, tmp_path):
    # the header is plain C: compile it with gcc and compare sizeof/offsetof with the ctypes mirror
    import subprocess
    from jspace_b200 import _lib as L
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "jspace_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(jsb_params), sizeof(jsb_run_opts),'
        ' sizeof(jsb_event), sizeof(jsb_summary), sizeof(jsb_clone), sizeof(jsb_snapshot), sizeof(jsb_relation),'
        ' offsetof(jsb_params, t_bottleneck), offsetof(jsb_params, root_rate), offsetof(jsb_run_opts, log_pool_bytes));'
        'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(L.Params), C.sizeof(L.RunOpts), L.EVENT_DTYPE.itemsize, L.SUMMARY_DTYPE.itemsize,
            L.CLONE_DTYPE.itemsize, L.SNAPSHOT_DTYPE.itemsize, L.RELATION_DTYPE.itemsize,
            L.Params.t_bottleneck.offset, L.Params.root_rate.offset, L.RunOpts.log_pool_bytes.offset]
    assert got == want
    assert got[2:7] == [32, 64, 16, 32, 24]


def test_header_constants_match_the_python_mirror(jsb):
    # every numeric #define JSB_* of the header has the same value in jspace_b200._lib (name without the prefix)
    from jspace_b200 import _lib as L
    txt = open(os.path.join(ROOT, "include", "jspace_b200.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define JSB_([A-Z0-9_]+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))u?\)?\s", txt)}
    assert len(defs) >= 20
    checked = 0
    for name, val in defs.items():
        if hasattr(L, name):
            assert getattr(L, name) == val, name
            checked += 1
    for prefix in ("QUIRK_", "OUT_", "MODE_", "ST_"):
        for name in defs:
            if name.startswith(prefix):
                assert hasattr(L, name), name
    assert checked >= 15


def test_graph_builders_match_oracle(jsb, oracle):
    for dim, row, col in [(2, 7, 5), (2, 100, 100), (3, 25, 25), (3, 350, 350), (5, 12, 12), (2, 1, 9)]:
        gj = jsb.Graph.lattice(row, col, dim)
        go = oracle.Graph.lattice(dim, row, col)
        assert gj.nv == go.nv
        rng = np.random.default_rng(dim * 1000 + row)
        for v in {1, gj.nv, *rng.integers(1, gj.nv + 1, 40).tolist()}:
            assert gj.neighbors(v).tolist() == go.neighbors(v).tolist(), (dim, row, col, v)
    # seed vertex rule of src/J_Space.jl:131-133
    assert jsb.Graph.lattice(100, 100, 2).seed_candidates().tolist() == [5051]
    assert jsb.Graph.lattice(200, 200, 2).seed_candidates().tolist() == [20101]
    assert jsb.Graph.lattice(350, 350, 3).seed_candidates().tolist() == [61426]
    assert jsb.Graph.lattice(25, 25, 3).seed_candidates().tolist() == [365]  # 9^3 = 729 vertices (odd)


def test_adjacency_file_and_closeness(jsb, oracle):
    path = os.path.join(ROOT, "tests", "golden", "Example_adj_matrix.txt")
    gj, go = jsb.Graph.dense_file(path), oracle.Graph.dense_file(path)
    assert gj.nv == 10 and gj.ne == 9
    for v in range(1, 11):
        assert gj.neighbors(v).tolist() == go.neighbors(v).tolist()
    assert gj.seed_candidates().tolist() == go.closeness_argmax().tolist() == [1, 2]
    from graphs import random_geometric_csr
    rowptr, colidx = random_geometric_csr(600, seed=3)
    gj, go = jsb.Graph.csr(rowptr, colidx), oracle.Graph.csr(rowptr, colidx)
    assert gj.seed_candidates().tolist() == go.closeness_argmax().tolist()


def test_errors_are_reported_not_thrown(jsb, tmp_path):
    with pytest.raises(jsb.JSpaceError, match="JSB_EINVAL"):
        jsb.Graph.lattice(0, 5, 2)
    with pytest.raises(jsb.JSpaceError, match="JSB_EIO"):
        jsb.Graph.dense_file(str(tmp_path / "missing.txt"))
    bad = tmp_path / "bad.txt"
    bad.write_text("0 1 0\n1 0 1\n")
    with pytest.raises(jsb.JSpaceError, match="square"):
        jsb.Graph.dense_file(str(bad))
    asym = tmp_path / "asym.txt"
    asym.write_text("0 1\n0 0\n")
    with pytest.raises(jsb.JSpaceError, match="symmetric"):
        jsb.Graph.dense_file(str(asym))
    with pytest.raises(jsb.JSpaceError, match="ascending"):
        jsb.Graph.csr([0, 2, 3, 4], [3, 2, 1, 1])
    with pytest.raises(jsb.JSpaceError, match="self-loop"):
        jsb.Graph.csr([0, 1, 1], [1])
    g = jsb.Graph.lattice(10, 10, 2)
    with pytest.raises(jsb.JSpaceError, match="ratio_bottleneck"):
        jsb.run_ensemble(g, jsb.Point(t_bottleneck=[1.0], ratio_bottleneck=[1.5]), 1)
    with pytest.raises(jsb.JSpaceError, match="n_start_cells"):
        jsb.run_ensemble(g, jsb.Point(n_start_cells=101), 1)
    with pytest.raises(jsb.JSpaceError, match="method"):
        jsb.run_ensemble(g, jsb.Point(method=3), 1)
    with pytest.raises(jsb.JSpaceError, match="replicate range"):
        jsb.run_ensemble(g, jsb.Point(), 4, g_begin=3, g_end=9)


def test_no_cpu_fallback(jsb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = jsb.Graph.lattice(10, 10, 2)
    with pytest.raises(jsb.JSpaceError, match="JSB_ENODEVICE"):
        jsb.run_ensemble(g, jsb.Point(), 1)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "j-space.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "jsb_kernels.cu" and "import oracle" not in txt, (dirpath, f)
