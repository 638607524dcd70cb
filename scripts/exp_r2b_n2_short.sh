#!/bin/bash
# round 2 (second session): the multi-rank flow of bench.py on two GPUs, shortened (16 groups per rank, 256 C5 groups)
set -x
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 6 --warmup 3 --groups 16 --c5-groups 256 ) > gpurun_out/r2b_bench_n2_short.json 2> gpurun_out/r2b_bench_n2_short.err
tail -6 gpurun_out/r2b_bench_n2_short.err
cut -c1-3000 gpurun_out/r2b_bench_n2_short.json
