#!/bin/bash
# Final-build C2 launch list (ncu) for profiles/ and profiles/traffic.json; run through gpurun.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
O=gpurun_out
python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu-baseline --no-fast > $O/r2f_bench_C2.json 2> $O/r2f_bench_C2.err || exit 1
ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/r2f_bench_C2_launches.csv \
    python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu-baseline --no-fast > $O/r2f_ncu_C2.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_reference.json 2>&1
