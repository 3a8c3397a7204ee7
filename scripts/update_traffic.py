#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list (CSV written by
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv --log-file X bench.py ...).
usage: update_traffic.py launches.csv WORKLOAD REPS  (run from the repo root, on the build the list was taken with)

Takes the LAST main ssa_kernel launch of the list (the timed step of `--steps 1 --warmup 1`; main launches are the
long ones, the pilot launches the short ones), adds its DRAM bytes read + written, and records the value under "WORKLOAD:REPS" together with the hash of
the kernel sources, so that bench.py only quotes it for the build it was measured on."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_sha16 only)


def main():
    path, workload, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = {}
    for r in rows:
        if "ssa_kernel" not in r.get("Kernel Name", ""):
            continue
        k = r["ID"]
        per.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}

    def dur(m):
        v, u = m["gpu__time_duration.sum"]
        return v * tscale[u]
    longest = max(dur(m) for m in per.values())
    mains = [k for k in per if dur(per[k]) > 0.5 * longest]      # main launches (the pilot launches are the short ones)
    main_id = max(mains, key=lambda k: int(k))                    # the last one = the timed step of the bench command
    m = per[main_id]
    rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    out = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(out))
    except Exception:
        t = {}
    sha = bench.source_sha16()
    if t.get("src_sha16") != sha:
        t = {"src_sha16": sha, "entries": {}, "detail": {}}
    t["entries"][f"{workload}:{reps}"] = rd + wr
    t["detail"][f"{workload}:{reps}"] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "kernel_s_under_ncu": dur(m),
                                         "source": os.path.basename(path)}
    json.dump(t, open(out, "w"), indent=1)
    print(f"{workload}:{reps} read {rd / 1e9:.2f} GB write {wr / 1e9:.2f} GB duration {dur(m):.3f} s -> {out}")


if __name__ == "__main__":
    main()
