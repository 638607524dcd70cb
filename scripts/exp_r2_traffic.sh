set -x
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:pp_search_kernel -c 1 --csv --log-file gpurun_out/r2_search_traffic.csv python scripts/run_bulk.py --slots 2368 --max-pops 60000 > gpurun_out/r2_bulk_traffic.log 2>&1
tail -2 gpurun_out/r2_bulk_traffic.log; cat gpurun_out/r2_search_traffic.csv | tail -8
