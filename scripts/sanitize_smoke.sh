#!/bin/bash
# compute-sanitizer over the smoke scenario (the reference's own test query in EXACT mode + K-POP(32)): memcheck and racecheck
# (shared-memory hazards of the warp-cooperative code).  Logs go to gpurun_out/; copy the ones to keep into profiles/.
set -x
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
  echo "exit $?" >> gpurun_out/r2_sanitizer_$TOOL.log
  tail -6 gpurun_out/r2_sanitizer_$TOOL.log
done
