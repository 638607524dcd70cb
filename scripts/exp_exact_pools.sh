#!/usr/bin/env bash
# Round-2 experiment (DESIGN.md sections 7 and 13): does the EXACT-mode C4 step get shorter when the first-pass pools are sized
# for the batch's longest query (no re-run) at the price of fewer resident queries?  Baseline = library defaults (2 368 slots,
# 131 072 closed states, re-run of the 116 long queries in 8x pools: 11.4 s + 28.1 s).  Run under gpurun, ~2 min per line:
#   gpurun --timeout 900 -- 'bash scripts/exp_exact_pools.sh > gpurun_out/exp_exact_pools.jsonl 2>gpurun_out/exp_exact_pools.err'
# Each line is bench.py's JSON line; compare ms_per_step, config.slots and config.pools.retried_queries.
set -u
common="--no-kpop --no-cpu-baseline --steps 1 --warmup 1"
python bench.py $common                                                                                   # measured default
for slots in 592 1024; do
  python bench.py $common --max-expansions 1048576 --max-open 524288 --max-open2d 65536 --exact-slots $slots
done
python bench.py $common --exact-slots 592                                                                 # default pools, 1 CTA / SM: the SM's saturation curve
