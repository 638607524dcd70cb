#!/bin/bash
# round 2 (second session): per-instruction execution counts and warp-stall samples of pp_search_kernel (ncu source page) at bench
# occupancy, on a SHORT launch (every query of the C4 batch stopped after 1500 expansions) in a context whose memory budget is
# 16 GB: ncu's kernel replay saves and restores every allocated device byte per pass -- with the default budget (75 % of 180 GB)
# the first two attempts never finished a pass.  Read back with
#   ncu -i gpurun_out/r2b_search_src_2368.ncu-rep --page source --csv > src.csv ; python scripts/ncu_lines.py src.csv all.sass <kernel>
set -x
mkdir -p gpurun_out
SEC="--section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy"
timeout 200 ncu $SEC --import-source on --clock-control none -k regex:pp_search_kernel -c 1 -f -o gpurun_out/r2b_search_src_2368 \
    python scripts/run_capped.py --slots 2368 --cap 1500 --budget-gb 16 > gpurun_out/r2b_ncu_src_2368.log 2>&1
tail -3 gpurun_out/r2b_ncu_src_2368.log
ls -la gpurun_out/*.ncu-rep
