#!/bin/bash
# round 2 (second session): GPU parity suite + bulk / lone-warp throughput of the search kernel after the instruction-footprint work
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2b_gputests.log
python scripts/exp_r2_slots.py --slots 148,592,1184,2368 > gpurun_out/r2b_slots.log 2>&1
tail -8 gpurun_out/r2b_slots.log
cat gpurun_out/r2b_gputests.log
