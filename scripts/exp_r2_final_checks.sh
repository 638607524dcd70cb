set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pp_search\|pp_kpop\|pp_field2d -c 400 --csv --log-file gpurun_out/r2_bench_search_launches.csv \
    python bench.py --steps 2 --warmup 1 --lanes 2 --e2e-steps 1 --no-c5 --no-cpu-baseline --no-blocks > gpurun_out/r2_bench_short_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:pp_map -c 80 --csv \
    --log-file gpurun_out/r2_map_launches.csv python scripts/bench_kernels.py > gpurun_out/r2_bench_kernels_ncu2.log 2>&1
python scripts/bench_kernels.py 2>&1 | grep -E "pp_map|field2d" | cut -c1-200
