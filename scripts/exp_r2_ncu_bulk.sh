set -x
python scripts/run_bulk.py --slots 2368 --max-pops 20000 > gpurun_out/r2_bulk_plain.log 2>&1
cat gpurun_out/r2_bulk_plain.log
ncu --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section InstructionStats --clock-control none -k regex:pp_search_kernel -c 1 --csv --log-file gpurun_out/r2_search_bulk_sections.csv python scripts/run_bulk.py --slots 2368 --max-pops 20000 > gpurun_out/r2_bulk_ncu.log 2>&1
tail -3 gpurun_out/r2_bulk_ncu.log
