"""The C ABI driven from C: tests/c_abi_driver.c performs the Julia wrapper's call sequence
(j-space.jl_b200/julia/JSpaceB200.jl) with plain C types, in two separate processes whose results must
agree (compute-sanitizer is closed on the GPU pool, so there is no memcheck leg)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "j-space.jl_b200", "lib")
SRC = os.path.join(ROOT, "tests", "c_abi_driver.c")


def build(tmp_path):
    exe = str(tmp_path / "c_abi_driver")
    subprocess.run(["gcc", "-O1", "-Wall", "-Wextra", "-std=c11", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", LIBDIR, "-ljspace_b200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def test_c_driver_compiles_and_links_without_a_gpu(tmp_path):
    """CPU: the header is valid C11 and every symbol the driver uses resolves against the library."""
    build(tmp_path)


@pytest.mark.gpu
def test_c_driver_call_sequence(tmp_path):
    exe = build(tmp_path)
    a = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0, a.stderr
    b = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert a.stdout.startswith("C-ABI-DRIVER OK") and a.stdout == b.stdout  # deterministic across processes
