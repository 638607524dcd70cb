#!/bin/bash
# round 2 (second session): ncu launch list of the bench command, search phase (reduced step count: under ncu every launch is
# serialised; single-pass metric, no kernel replay)
set -x
mkdir -p gpurun_out
timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pp_search\|pp_kpop\|pp_field2d -c 400 --csv \
    --log-file gpurun_out/r2b_bench_search_launches.csv \
    python bench.py --steps 2 --warmup 3 --lanes 2 --e2e-steps 1 --no-c5 --no-cpu-baseline --no-blocks > gpurun_out/r2b_bench_short_ncu.log 2>&1
tail -3 gpurun_out/r2b_bench_short_ncu.log | cut -c1-400
grep -c pp_search_kernel gpurun_out/r2b_bench_search_launches.csv
tail -12 gpurun_out/r2b_bench_search_launches.csv | cut -c1-220
