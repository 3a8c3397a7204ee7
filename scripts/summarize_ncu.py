#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + the source lines with the
most warp-stall samples.  usage: summarize_ncu.py report.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_fp64.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for r in rows[2:]:
    print(f"== kernel: {r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
    for h, u, v in zip(hdr, units, r):
        if h in KEYS or ("stalled" in h and "per_warp_active.pct" in h and v and float(v) > 2.0) or \
                ("smsp__average_warp" in h and "per_issue_active" in h and v and float(v) > 0.5):
            print(f"{h:75s} {v:>20s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        agg.append((cur, int(r[0]), r[1].strip()[:95], int(r[4]), int(r[7])))
    except ValueError:
        pass
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f"\n== source lines by warp-stall samples (total samples {ts}, total warp instructions {ti})")
for a in sorted(agg, key=lambda a: -a[3])[:40]:
    print(f"{a[0][:16]:16s}:{a[1]:5d} {100 * a[3] / ts:5.1f}% samples {100 * a[4] / ti:5.1f}% inst | {a[2]}")
