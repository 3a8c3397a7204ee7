"""bench.py contract pieces that do not need a GPU: the reference arm (CPU oracle pool) prints one JSON line with
the keys the driver reads, and the traffic figure is only quoted for the build it was measured on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, JSB_REF_STEP_S="0.4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "C1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["workload"] == "C1" and line["scaling"] == "weak"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["busy_fraction"] > 0.9 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", JSB_REF_STEP_S="0.2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_traffic_figure_is_keyed_to_the_build():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    got = bench.measured_traffic("C2", 4096)
    if t["src_sha16"] == bench.source_sha16():
        assert got == t["entries"]["C2:4096"] and got > 5e10
    else:
        assert got is None
    assert bench.measured_traffic("C2", 123) is None
