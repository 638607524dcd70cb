#!/bin/bash
# Round-2 profile records (run on a B200 through gpurun; outputs in gpurun_out/, copy the ones to keep into profiles/).
set -x
# 1. the bench command, un-profiled, then its ncu launch list (reduced step count: under ncu every launch is serialised)
python bench.py --steps 2 --warmup 1 --lanes 2 --e2e-steps 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --lanes 2 --e2e-steps 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_bench_short_ncu.log 2>&1
# 2. map / field kernels alone: C2 round (fused and two-call), lane lines, C3 fields
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:pp_map\|pp_field2d\|pp_dubins_field -c 60 --csv \
    --log-file gpurun_out/r2_map_field_launches.csv python scripts/bench_kernels.py > gpurun_out/r2_bench_kernels_ncu.log 2>&1
tail -3 gpurun_out/r2_bench_kernels_ncu.log
