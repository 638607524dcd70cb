set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import path_planning_pkg_b200 as pp, scenarios as S, bench
# C2 / C3 blocks only
print(bench.block_c2(None, pp, 0, 6546.6))
print(bench.block_c3(None, pp, 0, 6546.6))
# single-query latency anatomy on C1
sc = S.c1_scenario(3)
P = pp.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
ctx = pp.Context(P, num_groups=1)
S.build_map(ctx, sc)
q = ctx.make_queries(np.array([sc["queries"][0]]), [0])
o = ctx.make_opts(path_cap=2048, max_slots=1)
ctx.find_path_batch(q, o)
for rep in range(3):
    t0 = time.perf_counter(); ctx.batch_upload(q, o); t1 = time.perf_counter(); ms = ctx.batch_run(); t2 = time.perf_counter(); r = ctx.batch_fetch(want_paths=True); t3 = time.perf_counter()
    print(f"C1 single query: pops {int(r[0]['n_pops'][0])} upload {1e3*(t1-t0):.3f} ms, run {1e3*(t2-t1):.3f} ms (kernel {ms:.3f} ms), fetch {1e3*(t3-t2):.3f} ms")
PY
bash scripts/sanitize_smoke.sh
bash scripts/exp_r2_traffic.sh
