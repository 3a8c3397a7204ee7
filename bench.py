#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 SSA path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA path (libjspace_b200.so)
  python bench.py --impl reference --steps K --warmup W    # reference arm: CPU oracle port

Workload (config.workload = "C2"): BASELINE config[1], the one the metric is quoted on —
2D lattice 200x200, Parameters.toml dynamics ("contact", Tf=200, birth .2, death .01,
migration .01, driver rate .01, advantage N(.4,.1), one excision t=50 ratio .8), wild-type seed
cell at the lattice centre, 4096 independent replicates per GPU (weak scaling: rank r simulates
global replicates r*4096 .. r*4096+4095, replicate g always uses Philox stream g).

A "step" is one pass of the whole ensemble with seed = base_seed + step.  The unit of work is one
iteration of the reference's while loop (src/J_Space.jl:687), null events included
(SURVEY.md §8d).  `value` = iterations of all ranks / max-over-ranks device time of the
simulation kernel (CUDA events on the launching stream, inside the library).  `e2e` = the same
metric through the public C-ABI call with host buffers: jsb_run with the event log, the final
lattice and the clone tables copied back to host memory, wall-clocked on the host.

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (the CPU
restatement of the reference, timed on the host cores; the reference itself is Julia and there
is no Julia in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "j-space.jl_b200"))

WORKLOADS = {
    # name: (dim, row, col, replicates per GPU, point kwargs)
    "C2": (2, 200, 200, 4096, dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01,
                                    adv_mean=0.4, adv_std=0.1, rule="contact", t_bottleneck=[50.0],
                                    ratio_bottleneck=[0.8])),
    "C1": (2, 100, 100, 4096, dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01,
                                    adv_mean=0.4, adv_std=0.1, rule="contact", t_bottleneck=[50.0],
                                    ratio_bottleneck=[0.8])),
}
BASE_SEED = 1234
# algorithmic bytes per loop iteration, SURVEY.md §8(d)
A_NULL, A_BIRTH, A_DEATH = 6.0, 78.0, 46.0
# dram__bytes_read.sum + dram__bytes_write.sum of the main ssa_kernel launch of the timed step of workload C2
# (4096 replicates), from the ncu pass committed as profiles/r1_bench_c2_launches.csv (launch ID 3: 48.64 GB read +
# 39.74 GB written; the 27 ms pilot launch before it moves another 0.8 GB)
MEASURED_TRAFFIC = {("C2", 4096): 88381294336}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(summaries):
    it = float(summaries["n_iterations"].sum())
    births = float(summaries["n_births"].sum())
    deaths = float(summaries["n_deaths"].sum() + summaries["n_migrations"].sum() + summaries["n_excised"].sum()
                   + summaries["n_displaced"].sum())
    return A_NULL * it + A_BIRTH * births + A_DEATH * deaths


def oracle_threads(workload, reps, seed, n_threads, faithful=True, max_seconds=None):
    """One oracle replicate per host thread (ctypes releases the GIL).  Returns (iterations, seconds,
    replicates done)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O  # test infrastructure: allowed here (cpu_baseline / reference arm only)
    dim, row, col, _, kw = WORKLOADS[workload]
    g = O.Graph.lattice(dim, row, col)
    todo = list(range(reps))
    lock = threading.Lock()
    total = [0, 0]
    t0 = time.perf_counter()

    def work():
        while True:
            with lock:
                if not todo or (max_seconds and time.perf_counter() - t0 > max_seconds):
                    return
                s = todo.pop(0)
            r = O.Run(g, O.make_params(**kw), seed, s, faithful=faithful, want=())
            with lock:
                total[0] += int(r.summary["n_iterations"])
                total[1] += 1

    th = [threading.Thread(target=work) for _ in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return total[0], time.perf_counter() - t0, total[1]


def run_reference(args):
    """--impl reference: the CPU restatement of the reference loop (oracle port, faithful O(n)
    scans) on all host cores, one replicate per thread, on the same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    wl = args.workload
    reps_per_step = cores  # bounded sample: one replicate per host thread per step
    for w in range(args.warmup):
        oracle_threads(wl, min(reps_per_step, 2), BASE_SEED + 1000 + w, cores)
    iters = 0
    secs = 0.0
    done = 0
    for s in range(args.steps):
        it, dt, n = oracle_threads(wl, reps_per_step, BASE_SEED + s, cores)
        iters += it
        secs += dt
        done += n
    value = iters / secs if secs > 0 else 0.0
    sample = (f"{done} full replicates of workload {wl} (streams 0..{reps_per_step - 1} per step), one per host "
              f"thread, oracle port with the reference's O(n) membership scans")
    line = {
        "impl": "reference", "metric": "SSA loop iterations per second (null events included)", "value": value,
        "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
        "config": {"workload": wl, "replicates_per_step": reps_per_step},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--reps", type=int, default=0, help="replicates per GPU (default: the workload's 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fast", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import jspace_b200 as J
    from jspace_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: bench.py has no CPU fallback for the product path"}))
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dim, row, col, reps_default, kw = WORKLOADS[args.workload]
    reps = args.reps or reps_default
    graph = J.Graph.lattice(row, col, dim)
    point = J.Point(**kw)
    g0 = rank * reps  # weak scaling: every rank gets its own `reps` streams
    total_reps = world * reps

    def step(seed, flags=0, pool=0):
        return J.run_ensemble(graph, point, total_reps, seed=seed, flags=flags, g_begin=g0, g_end=g0 + reps,
                              device=local_rank, log_pool_bytes=pool)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(BASE_SEED + 1000 + w).close()

    # ---- timed: device time of the simulation kernel, inputs resident --------------------------
    barrier()
    kernel_ms = 0.0
    iters = 0
    alg_bytes = 0.0
    launches = 0
    eff = 0
    hist = np.zeros(64, dtype=np.int64)
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        for s in range(args.steps):
            r = step(BASE_SEED + s)
            kernel_ms += r.kernel_ms
            iters += int(r.summaries["n_iterations"].sum())
            eff += int(r.summaries["n_effective"].sum())
            alg_bytes += algorithmic_bytes(r.summaries)
            launches += r.launches
            hist += r.histogram(0, 0).astype(np.int64)
            r.close()
        barrier()
        wall = time.perf_counter() - wall0
    t = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(iters), float(eff), alg_bytes, float(launches)], dtype=torch.float64, device="cuda")
    h = torch.from_numpy(hist).cuda()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks of the device time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(h, op=dist.ReduceOp.SUM)   # the ensemble summary histogram gather (NCCL)
    max_ms = float(t.item())
    all_iters, all_eff, all_bytes, all_launches = [float(x) for x in tot.tolist()]
    value = all_iters / (max_ms / 1e3)

    # ---- e2e: public C-ABI call, host buffers, every output of the reference copied back -------
    e2e = None
    if not args.no_e2e:
        flags = L.OUT_EVENTS | L.OUT_LATTICE
        pool = 14 << 30
        step(BASE_SEED + 2000, flags, pool).close()  # warm the pinned staging buffer
        barrier()
        e_iters = 0
        d2h = 0
        e_wall0 = time.perf_counter()
        for s in range(args.steps):
            r = step(BASE_SEED + s, flags, pool)
            e_iters += int(r.summaries["n_iterations"].sum())
            d2h += (int(r.summaries["n_events"].sum()) * 32 + r.count * 64 + int(r.summaries["n_clones"].sum()) * 16
                    + r.count * graph.nv * 6)
            assert int((r.summaries["status"] == L.ST_LOG_OVERFLOW).sum()) == 0, "event-log pool too small"
            r.close()
        barrier()
        e_wall = time.perf_counter() - e_wall0
        et = torch.tensor([e_wall], dtype=torch.float64, device="cuda")
        ei = torch.tensor([float(e_iters), float(d2h)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(ei, op=dist.ReduceOp.SUM)
        e2e = {"value": ei[0].item() / et.item(), "unit": "iterations/s",
               "h2d_bytes_per_step": int(world * (256 + 8 * 4)),
               "d2h_bytes_per_step": int(ei[1].item() / max(1, args.steps)),
               "what": "jsb_run with JSB_OUT_EVENTS|JSB_OUT_LATTICE: event logs, final lattices, clone tables and "
                       "summaries of every replicate copied to host memory; host wall clock"}

    # ---- extra: the rejection-free mode on the same ensemble, one pass --------------
    fast = None
    if not args.no_fast:
        barrier()
        rf = J.run_ensemble(graph, point, total_reps, seed=BASE_SEED, flags=L.MODE_REJECTION_FREE, g_begin=g0,
                            g_end=g0 + reps, device=local_rank)
        ft = torch.tensor([rf.kernel_ms], dtype=torch.float64, device="cuda")
        fe = torch.tensor([float(rf.summaries["n_iterations"].sum())], dtype=torch.float64, device="cuda")
        launches += rf.launches
        rf.close()
        if dist is not None:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
            dist.all_reduce(fe, op=dist.ReduceOp.SUM)
        fast = {"replicates_per_s": world * reps / (ft.item() / 1e3), "state_changing_events_per_s": fe.item() / (ft.item() / 1e3),
                "ms": ft.item(), "note": "JSB_MODE_REJECTION_FREE: same law, null events not simulated; KS-validated, "
                                         "not event-for-event comparable"}

    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores, bounded sample --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_rep = cores
        it, dt, done = oracle_threads(args.workload, n_rep, BASE_SEED, cores)
        cpu = {"value": it / dt, "unit": "iterations/s", "cores": cores, "kind": "port",
               "sample": f"{done} full replicates of workload {args.workload} (streams 0..{n_rep - 1}, seed "
                         f"{BASE_SEED}), one per host thread, oracle port with the reference's O(n) membership scans; "
                         f"{dt:.1f} s"}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = all_bytes / (max_ms / 1e3) / 1e9 / world  # per GPU, per launch == per step
        line = {
            "metric": "SSA loop iterations per second (null events included)",
            "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
            "config": {"workload": args.workload, "lattice": [row, col], "dim": dim, "replicates_per_gpu": reps,
                       "parallelism": f"replicas x{world}", "l2": "per-step working set 2.5 GB >> 126 MB L2",
                       "params": {k: v for k, v in kw.items()}},
            "replicates_per_s": all_launches and (world * reps * args.steps) / (max_ms / 1e3),
            "effective_events_per_s": all_eff / (max_ms / 1e3),
            "wall_s": wall,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": MEASURED_TRAFFIC.get((args.workload, reps)), "peak_source": peak_src,
                         "note": "latency-bound SSA: algorithmic bytes 6 B/iteration + 78 B/birth + 46 B/death "
                                 "(SURVEY 8d); see profiles/ for issue-slot utilisation and stall reasons"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(all_launches),
            "clocks": clocks.summary(),
            "ensemble_histogram_n_alive": h.tolist()[:8],
            "rejection_free_mode": fast,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
