"""The C ABI driven from C: tests/c_abi_driver.c performs the Julia wrapper's call sequence
(j-space.jl_b200/julia/JSpaceB200.jl) with plain C types, twice, and once under compute-sanitizer
(memcheck: out-of-bounds / misaligned device accesses, leaks of device allocations at exit)."""
import os
import shutil
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


@pytest.mark.gpu
def test_c_driver_under_compute_sanitizer(tmp_path):
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(cs):
        pytest.skip("compute-sanitizer not installed")
    exe = build(tmp_path)
    r = subprocess.run([cs, "--tool", "memcheck", "--leak-check", "full", "--error-exitcode", "9", exe],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "C-ABI-DRIVER OK" in r.stdout
