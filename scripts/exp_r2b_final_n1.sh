#!/bin/bash
# round 2 (second session): GPU test suite + the driver's bench command at N = 1 on the final sources
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2b_gputests_final.log
cat gpurun_out/r2b_gputests_final.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_n1_steps20.json 2> gpurun_out/r2b_bench_n1_steps20.err
tail -5 gpurun_out/r2b_bench_n1_steps20.err
cut -c1-1500 gpurun_out/r2b_bench_n1_steps20.json
