#!/usr/bin/env python
"""Development helper: ONE launch of the EXACT search kernel on the C4 batch with every query stopped after --cap expansions
(no retry pass: the launch is not waited for through pp_batch_wait).  A short launch at bench occupancy for instrumented ncu
passes (SourceCounters multiplies the run time of this kernel by two orders of magnitude)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import path_planning_pkg_b200 as pp  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=2368)
ap.add_argument("--cap", type=int, default=1500)
ap.add_argument("--groups", type=int, default=64)
ap.add_argument("--budget-gb", type=float, default=0.0, help="memory budget of the context (ncu kernel replay saves and restores every allocated byte per pass)")
a = ap.parse_args()
P = pp.make_params(grid_size=512, resolution=0.2)
ctx = pp.Context(P, num_groups=a.groups)
if a.budget_gb > 0:
    ctx.set_memory_budget(int(a.budget_gb * 1e9))
groups = bench.build_workload(a.groups, 64, 0)
bench.apply_groups(ctx, groups)
queries, qgroups, _ = bench.select_queries(ctx, groups)
q = ctx.make_queries(queries, qgroups)
ctx.batch_upload(q, ctx.make_opts(path_cap=2048, max_slots=a.slots, max_expansions=a.cap))
t0 = time.perf_counter()
ctx.batch_run_async()
ctx.sync()
print(f"{len(q)} queries capped at {a.cap} expansions, slots {a.slots}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
os._exit(0)      # no pp_batch_wait: the capped queries would be re-run with larger caps
