#!/bin/bash
# round 2 (second session): the e2e region of the final bench.py with 10 steps (what the driver's command runs at N = 1), on a
# shortened command (3 warm-up, 10 timed steps, no side blocks)
set -x
mkdir -p gpurun_out
( time timeout 380 python bench.py --gpus 1 --steps 10 --warmup 3 --lanes 20 --e2e-steps 10 --no-kpop --no-c5 --no-blocks --no-cpu-baseline ) \
    > gpurun_out/r2b_bench_e2e10.json 2> gpurun_out/r2b_bench_e2e10.err
tail -4 gpurun_out/r2b_bench_e2e10.err
cut -c1-2500 gpurun_out/r2b_bench_e2e10.json
