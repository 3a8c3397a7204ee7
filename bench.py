#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 SSA path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # CUDA path (libjspace_b200.so)
  python bench.py --impl reference --steps K --warmup W    # reference arm: CPU oracle port

Workloads (--workload, default C2 = BASELINE config[1], the one the metric is quoted on):
  C2  2D lattice 200x200, Parameters.toml dynamics ("contact", Tf=200, birth .2, death .01, migration .01,
      driver rate .01, advantage N(.4,.1), one excision t=50 ratio .8), wild-type seed cell at the lattice
      centre, 4096 independent replicates per GPU
  C1  the same on 100x100;  C3  50^3 lattice, linear 4-driver tree, hvoter, excision, 256 replicates;
  C4  ~100 k-node random geometric graph (CSR path), 1024 replicates;  C5  128-point parameter grid on 100x100,
      summary-only (default 2^17 replicates, --reps 1048576 for BASELINE's 2^20)
Weak scaling: rank r simulates its own `reps` replicates, replicate g always uses Philox stream g.

A "step" is one pass of the whole ensemble with seed = base_seed + step: ONE public C-ABI call (jsb_run) that
simulates every replicate and leaves every output the reference produces (event log, final lattice, clone
tables, summaries) in host memory.  The unit of work is one iteration of the reference's while loop
(src/J_Space.jl:687), null events included.  `value` = iterations of all ranks / max-over-ranks
device time of the simulation launches inside those calls (CUDA events on the launching stream; the kernel
writes the log and maintains cell ids, as the reference does).  `e2e` = the same iterations / max-over-ranks
host wall clock of the calls, copies included.

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (the CPU restatement of the
reference, timed on ALL host cores kept busy; the reference itself is Julia and there is no Julia in the image).
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

TREE_EDGES = [(0, 1), (1, 2), (2, 3)]   # utility/Tree_driver_example.txt (linear 4-driver tree)
TREE_RATES = [0.2, 0.5, 0.55, 0.6]      # utility/driver_advantage.txt
DEFAULT_DYN = dict(Tf=200.0, rate_birth=0.2, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, adv_mean=0.4,
                   adv_std=0.1, rule="contact")   # Parameters.toml:14-21
# BASELINE.json configs.  graph spec, replicates per GPU, sweep points (dicts of Point kwargs),
# outputs the e2e call copies back.
WORKLOADS = {
    "C1": dict(graph=("lattice", 2, 100, 100), reps=4096, outputs="log+lattice",
               points=[dict(DEFAULT_DYN, t_bottleneck=[50.0], ratio_bottleneck=[0.8])]),
    "C2": dict(graph=("lattice", 2, 200, 200), reps=4096, outputs="log+lattice",
               points=[dict(DEFAULT_DYN, t_bottleneck=[50.0], ratio_bottleneck=[0.8])]),
    "C3": dict(graph=("lattice", 3, 350, 350), reps=256, outputs="log+lattice",   # 50^3 sites, u32 site ids
               points=[dict(method=2, Tf=200.0, rate_death=0.01, rate_migration=0.01, mu_driver=0.01, rule="hvoter",
                            tree_edges=TREE_EDGES, node_rate=TREE_RATES, root_node=0, root_rate=0.2,
                            t_bottleneck=[100.0], ratio_bottleneck=[0.8])]),
    "C4": dict(graph=("rgg", 100000, 4321), reps=1024, outputs="log+lattice",     # CSR / HBM path
               points=[dict(DEFAULT_DYN)]),
    # 128-point birth/death/advantage grid on 100x100, summary-only; BASELINE size is 2^20 replicates (8192 per
    # point, 70 s per pass on one B200): the default here is 2^17 so that a default run finishes in minutes;
    # --reps 1048576 runs the full size
    "C5": dict(graph=("lattice", 2, 100, 100), reps=1 << 17, outputs="summaries",
               points=[dict(Tf=200.0, rate_birth=b, rate_death=d, rate_migration=0.01, mu_driver=1e-5, adv_mean=a,
                            adv_std=0.1, rule="contact")
                       for b in [0.1 + 0.5 * i / 7 for i in range(8)] for d in (0.005, 0.01, 0.02, 0.04)
                       for a in (0.1, 0.2, 0.4, 0.8)]),
}
BASE_SEED = 1234
# algorithmic bytes per loop iteration, SURVEY.md 8(d): member lookup 4 B + occupancy probe 2 B (+ 8 B rowptr pair
# + 4 B colidx on CSR graphs); + 78 B per birth, + 46 B per death / migration / excised / displaced cell
A_NULL = {"lattice": 6.0, "rgg": 18.0}
A_BIRTH, A_DEATH = 78.0, 46.0
METRIC = "SSA loop iterations per second (null events included)"


def source_sha16():
    """Identifies the build: hash of the kernel/host sources the library was compiled from."""
    import hashlib
    h = hashlib.sha256()
    for rel in ("j-space.jl_b200/csrc/jsb_kernels.cu", "j-space.jl_b200/csrc/jsb_api.cu",
                "j-space.jl_b200/csrc/jsb_device.cuh", "include/jspace_b200.h"):
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(workload, reps):
    """dram__bytes_read.sum + dram__bytes_write.sum of the main launch of one step, from the ncu capture
    recorded in profiles/traffic.json by scripts/update_traffic.py.  Only returned when the capture was taken on
    THIS build (source hash) and this workload size; otherwise null — a stale constant is worse than none."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("src_sha16") != source_sha16():
            return None
        return t.get("entries", {}).get(f"{workload}:{reps}")
    except Exception:
        return None


def config_of(workload, reps, world):
    """The `config` object: identical in both arms (the reference arm times the same workload)."""
    w = WORKLOADS[workload]
    return {"workload": workload, "graph": list(w["graph"]), "replicates_per_gpu": reps,
            "sweep_points": len(w["points"]), "parallelism": f"replicas x{world}",
            "l2": "per-step working set (>= 1 GB of per-replicate state) >> 126 MB L2",
            "params": w["points"][0] if len(w["points"]) == 1 else "128-point grid b x d x adv_mean (SURVEY 8d)"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_graph(J_or_O, spec, oracle=False):
    """Graph of a workload for the CUDA library (J = jspace_b200) or for the CPU oracle (O)."""
    if spec[0] == "lattice":
        _, dim, row, col = spec
        return (J_or_O.Graph.lattice(dim, row, col) if oracle else J_or_O.Graph.lattice(row, col, dim)), None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from graphs import geometric_centre, random_geometric_csr
    rowptr, colidx, pts = random_geometric_csr(spec[1], seed=spec[2], return_points=True)
    g = J_or_O.Graph.csr(rowptr, colidx)
    centre = geometric_centre(pts)  # stand-in for the max-closeness vertex
    if not oracle:
        g.set_seed_candidates([centre])
    return g, [centre]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device):
        """device: an ordinal or a comma-separated list of ordinals; None = do not sample (ranks other than 0: one
        nvidia-smi process watches every GPU of the job, eight of them polling the driver would perturb it)"""
        self.device = device
        self.rows = []
        self.proc = None

    def __enter__(self):
        if self.device is None:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "500"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        """Only the samples that arrived inside the timed region [t0, t1] count (the sampler is started before the
        warm-up so that the start-up of nvidia-smi, which holds driver locks, does not land in the timed region)."""
        self.t0, self.t1 = t0, t1

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
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        for ts, r in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.3):
                continue
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


def algorithmic_bytes(summaries, a_null):
    it = float(summaries["n_iterations"].sum())
    births = float(summaries["n_births"].sum())
    deaths = float(summaries["n_deaths"].sum() + summaries["n_migrations"].sum() + summaries["n_excised"].sum()
                   + summaries["n_displaced"].sum())
    return a_null * it + A_BIRTH * births + A_DEATH * deaths


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class OraclePool:
    """The reference's CPU path on ALL host cores for as long as the pool lives.

    One oracle replicate at a time per host thread (ctypes releases the GIL), streams 0, 1, 2, ... of the
    workload pulled from a shared counter, each replicate run to completion — replicate lengths are heavy
    tailed (4.5 % die at the first event, the heaviest takes 7x the mean), so a fixed batch joined on its
    slowest member leaves most cores idle; here a thread that finishes simply takes the next stream.  The
    oracle publishes its iteration count while it runs (orc_run_ctl), so a measurement is the difference of
    two marks of the live counters and never has to wait for (or throw away) the replicates in flight.
    `faithful=True`: the reference's own complexity (O(n) `pos in cs_alive` scans, O(n) filter!)."""

    def __init__(self, workload, seed, n_threads, faithful=True):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ctypes as C
        import numpy as np
        import oracle as O  # test infrastructure: allowed here (cpu_baseline / reference arm only)
        self.O, self.C, self.np = O, C, np
        w = WORKLOADS[workload]
        self.graph, self.cand = build_graph(O, w["graph"], oracle=True)
        self.points = [O.make_params(**kw) for kw in w["points"]]
        self.seed, self.n, self.faithful = seed, n_threads, faithful
        L = O.lib()
        L.orc_run_ctl.restype = C.c_void_p
        L.orc_run_ctl.argtypes = [C.c_void_p, C.POINTER(O.OrcParams), C.c_uint64, C.c_uint32, C.c_void_p, C.c_int64,
                                  C.c_int, C.c_void_p, C.c_void_p]
        self.progress = np.zeros(n_threads * 8, dtype=np.uint64)   # one cache line per thread
        self.stop = np.zeros(1, dtype=np.int32)
        self.done_it = [0] * n_threads
        self.done_reps = [0] * n_threads
        self.busy = [0.0] * n_threads
        self.t_enter = [None] * n_threads
        self.lock = threading.Lock()
        self.next_stream = 0
        self.threads = []

    def _work(self, k):
        O, C, L = self.O, self.C, self.O.lib()
        cand = self.np.ascontiguousarray(self.cand, dtype=self.np.int64) if self.cand else None
        prog_ptr = self.progress.ctypes.data + 64 * k
        while not self.stop[0]:
            with self.lock:
                s = self.next_stream
                self.next_stream += 1
                self.t_enter[k] = time.perf_counter()
            p, _keep = self.points[s % len(self.points)]
            h = L.orc_run_ctl(self.graph.h, C.byref(p), self.seed, s, cand.ctypes.data if cand is not None else None,
                              len(cand) if cand is not None else 0, 1 if self.faithful else 0, prog_ptr,
                              self.stop.ctypes.data)
            it = int(self.np.frombuffer(C.string_at(L.orc_summary(h), 64), dtype=O.SUMMARY_DTYPE)[0]["n_iterations"])
            L.orc_free(h)
            with self.lock:
                self.busy[k] += time.perf_counter() - self.t_enter[k]
                self.t_enter[k] = None
                self.done_it[k] += it
                self.done_reps[k] += 1
                self.progress[8 * k] = 0

    def start(self):
        self.threads = [threading.Thread(target=self._work, args=(k,), daemon=True) for k in range(self.n)]
        for t in self.threads:
            t.start()
        return self

    def mark(self):
        """(time, iterations so far incl. replicates in flight, replicates finished, busy thread-seconds)"""
        with self.lock:
            now = time.perf_counter()
            it = sum(self.done_it) + int(self.progress[::8].sum())
            busy = sum(self.busy) + sum(now - t for t in self.t_enter if t is not None)
            return now, it, sum(self.done_reps), busy

    def close(self):
        self.stop[0] = 1
        for t in self.threads:
            t.join()


def cpu_sample(workload, seconds, cores):
    """Rate of the saturated host over `seconds` (after a short ramp so that the threads are not all in the cheap
    first phase of their first replicate)."""
    pool = OraclePool(workload, BASE_SEED, cores).start()
    time.sleep(min(3.0, seconds / 4))
    t0, i0, r0, b0 = pool.mark()
    time.sleep(seconds)
    t1, i1, r1, b1 = pool.mark()
    pool.close()
    return (i1 - i0) / (t1 - t0), t1 - t0, r1 - r0, (b1 - b0) / (cores * (t1 - t0))


def run_reference(args):
    """--impl reference: the CPU restatement of the reference loop (oracle port with the reference's O(n)
    scans) on all host cores, same workload/config/metric.  A step is a fixed time box of the saturated pool
    (REF_STEP_S seconds of every core), so --steps 20 --warmup 5 ends within a few minutes; warm-up steps
    also carry the pool past its start-up transient (all threads in the cheap early phase of a replicate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    wl = args.workload
    reps = args.reps or WORKLOADS[wl]["reps"]
    step_s = float(os.environ.get("JSB_REF_STEP_S", "6.0"))
    pool = OraclePool(wl, BASE_SEED, cores).start()
    for _ in range(args.warmup):
        time.sleep(step_s)
    t0, i0, r0, b0 = pool.mark()
    for _ in range(args.steps):
        time.sleep(step_s)
    t1, i1, r1, b1 = pool.mark()
    pool.close()
    secs = t1 - t0
    value = (i1 - i0) / secs if secs > 0 else 0.0
    busy = (b1 - b0) / (cores * secs) if secs > 0 else 0.0
    sample = (f"{cores} host threads kept busy for {secs:.0f} s on streams 0.. of workload {wl} (seed {BASE_SEED}), "
              f"each replicate run to completion, {r1 - r0} finished inside the timed region, iterations of "
              f"replicates in flight counted as they happen; oracle port with the reference's O(n) membership "
              f"scans; busy fraction {busy:.3f}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
        "config": config_of(wl, reps, args.gpus),
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample,
                         "per_core": value / cores, "busy_fraction": busy},
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
    ap.add_argument("--reps", type=int, default=0, help="replicates per GPU (default: the workload's own)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="length of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    w = WORKLOADS[args.workload]
    reps = args.reps or w["reps"]
    graph, _ = build_graph(J, w["graph"])
    base_points = [J.Point(**kw) for kw in w["points"]]
    n_pts = len(base_points)
    if reps % n_pts:
        raise SystemExit(f"--reps must be a multiple of the {n_pts} sweep points")
    rpp = reps // n_pts
    # Weak scaling: every rank runs its own `reps` replicates with their own Philox streams.  Global numbering
    # g = point * rpp + r over `world` copies of the point list; rank k owns copy k, i.e. g in [k*reps, (k+1)*reps).
    points = base_points * world
    g0 = rank * reps
    flags = (L.OUT_EVENTS | L.OUT_LATTICE) if w["outputs"] == "log+lattice" else 0
    pool = (14 << 30) if flags else 0

    def step(seed, mode=0):
        return J.run_ensemble(graph, points, rpp, seed=seed, flags=flags | mode, g_begin=g0, g_end=g0 + reps,
                              device=local_rank, log_pool_bytes=pool)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(",".join(str(i) for i in range(world)) if rank == 0 else None)
    clocks.__enter__()
    for i in range(args.warmup):
        step(BASE_SEED + 1000 + i).close()

    # ---- timed region.  Every step is one public C-ABI call (jsb_run) that simulates the whole ensemble AND
    # brings every output the reference produces (event log `df`, final lattice, clone tables, summaries) back
    # into host memory.  `value` uses the CUDA-event time of the simulation launches inside the call (inputs
    # resident, outputs left in HBM); `e2e` uses the host wall clock of the same calls. ---------------------
    barrier()
    kernel_ms = 0.0
    call_s = 0.0   # host wall clock inside the public calls alone (the rest of the loop is this script's bookkeeping)
    iters = eff = launches = d2h = 0
    alg_bytes = 0.0
    a_null = A_NULL[w["graph"][0]]
    hist = np.zeros(64, dtype=np.int64)
    try:
        wall0 = time.perf_counter()
        for s in range(args.steps):
            tc0 = time.perf_counter()
            r = step(BASE_SEED + s)
            call_s += time.perf_counter() - tc0
            kernel_ms += r.kernel_ms
            sm = r.summaries
            iters += int(sm["n_iterations"].sum())
            eff += int(sm["n_effective"].sum())
            alg_bytes += algorithmic_bytes(sm, a_null)
            launches += r.launches
            d2h += r.count * 64 + int(sm["n_clones"].sum()) * 16
            if flags:
                d2h += int(sm["n_events"].sum()) * 32 + r.count * graph.nv * 6
                assert int((sm["status"] == L.ST_LOG_OVERFLOW).sum()) == 0, "event-log pool too small"
            hist += r.histogram(0, 0).astype(np.int64)
            r.close()
        barrier()
        wall = time.perf_counter() - wall0
        clocks.window(wall0, wall0 + wall)
    finally:
        clocks.__exit__(None, None, None)
    t = torch.tensor([kernel_ms, wall, call_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(iters), float(eff), alg_bytes, float(launches), float(d2h)], dtype=torch.float64,
                       device="cuda")
    h = torch.from_numpy(hist).cuda()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks of the device time / wall time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(h, op=dist.ReduceOp.SUM)   # the ensemble summary histogram gather (NCCL)
    max_ms, max_wall, max_call_ms = [float(x) for x in t.tolist()]
    all_iters, all_eff, all_bytes, all_launches, all_d2h = [float(x) for x in tot.tolist()]
    value = all_iters / (max_ms / 1e3)
    h2d = sum(256 + 8 * (len(kw.get("t_bottleneck", ())) * 2 + len(kw.get("node_rate", ())) + 2 * len(kw.get("tree_edges", ())))
              for kw in w["points"]) * world
    e2e = {"value": all_iters / max_wall, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(all_d2h / max(1, args.steps)),
           "inside_calls_s": max_call_ms / 1e3, "loop_wall_s": max_wall,
           "what": ("jsb_run with JSB_OUT_EVENTS|JSB_OUT_LATTICE: event logs, final lattices, clone tables and "
                    "summaries of every replicate in host memory when the call returns; host wall clock"
                    if flags else "jsb_run, summary-only sweep (the workload's definition): summaries and clone "
                                  "tables in host memory when the call returns; host wall clock")}

    # ---- extra: the rejection-free mode on the same ensemble, one pass --------------
    fast = None
    if not args.no_fast and args.workload != "C4":
        barrier()
        rf = J.run_ensemble(graph, points, rpp, seed=BASE_SEED, flags=L.MODE_REJECTION_FREE, g_begin=g0,
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

    # ---- CPU baseline (rank 0, N=1 only): oracle port on ALL host cores, time-boxed sample ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate, secs, done, busy = cpu_sample(args.workload, args.cpu_seconds, cores)
        cpu = {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port", "per_core": rate / cores,
               "busy_fraction": busy,
               "sample": f"{cores} host threads kept busy for {secs:.1f} s on streams 0.. of workload {args.workload} "
                         f"(seed {BASE_SEED}), each replicate run to completion ({done} finished), iterations of "
                         f"replicates in flight counted as they happen; oracle port with the reference's O(n) "
                         f"membership scans"}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = all_bytes / (max_ms / 1e3) / 1e9 / world  # per GPU, per launch == per step
        sector = None
        if w["graph"][0] == "rgg":  # sector-granular traffic of the CSR gathers, 4 x 32 B per iteration
            sector = 128.0 * all_iters / (max_ms / 1e3) / 1e9 / world
        line = {
            "metric": METRIC,
            "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
            "config": config_of(args.workload, reps, world),
            "replicates_per_s": (world * reps * args.steps) / (max_ms / 1e3),
            "effective_events_per_s": all_eff / (max_ms / 1e3),
            "wall_s": wall,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, reps), "peak_source": peak_src,
                         "achieved_sector_granular": sector,
                         "note": f"latency-bound SSA: algorithmic bytes {a_null:.0f} B/iteration + 78 B/birth + 46 B/death "
                                 "(SURVEY 8d); see profiles/ for issue-slot utilisation and stall reasons"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(all_launches),
            "clocks": clocks.summary(),
            "ensemble_histogram_n_alive": h.tolist()[:8],
            "rejection_free_mode": fast,
            "build": source_sha16(),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
