set -x
for V in w1 w2; do
  PP_B200_LIB=$PWD/path_planning_pkg_b200/lib/libpp_b200_$V.so python scripts/exp_r2_slots.py --slots 592,2368 2>&1 | grep -E "slots|full"
done
PP_B200_LIB=$PWD/path_planning_pkg_b200/lib/libpp_b200_w1.so python bench.py --steps 4 --warmup 2 --no-c5 --no-blocks --no-kpop --no-cpu-baseline > gpurun_out/r2_bench_w1.json 2> gpurun_out/r2_bench_w1.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_w1.json') if l.startswith('{')][0])
print('w1 bench: value',d['value'],'ms_per_step',d['ms_per_step'],'latency',d['batch_latency_ms'],'e2e',d['e2e']['value'])"
