#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; outputs under gpurun_out/).
# Every ncu command is preceded by the same command without ncu (must exit 0).
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
O=gpurun_out
for W in C2 C4 C3 C5; do
  EXTRA=""; [ $W = C5 ] && EXTRA="--reps 32768"
  python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --no-fast $EXTRA > $O/r2_bench_$W.json 2> $O/r2_bench_$W.err || exit 1
  ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/r2_bench_${W}_launches.csv \
      python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --no-fast $EXTRA > $O/r2_ncu_$W.log 2>&1
done
# full-set captures on short workloads (ncu replays a kernel ~40 times)
python scripts/prof_throughput.py 7 > $O/r2_thr7.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -c 1 -o $O/r2_throughput_7 -f python scripts/prof_throughput.py 7 > $O/r2_ncu_thr7.log 2>&1
python bench.py --workload C4 --reps 296 --steps 1 --warmup 0 --no-cpu-baseline --no-fast > $O/r2_c4_short.json 2>&1 || exit 1
JSB_NO_PILOT=1 ncu --set full --clock-control none --import-source on -k regex:ssa_kernel -c 1 -o $O/r2_c4_csr_u32 -f python bench.py --workload C4 --reps 296 --steps 1 --warmup 0 --no-cpu-baseline --no-fast > $O/r2_ncu_c4.log 2>&1
ls -la $O
