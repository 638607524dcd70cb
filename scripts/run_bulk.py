#!/usr/bin/env python
"""Development helper: ONE launch of the EXACT search kernel on the C4 batch without its long tail (for ncu at bench occupancy).
The per-query expansion counts come from a cached first run (gpurun_out/c4_pops.npy) or are computed once."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import path_planning_pkg_b200 as pp  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=2368)
ap.add_argument("--max-pops", type=int, default=20000)
ap.add_argument("--groups", type=int, default=64)
a = ap.parse_args()
P = pp.make_params(grid_size=512, resolution=0.2)
ctx = pp.Context(P, num_groups=a.groups)
groups = bench.build_workload(a.groups, 64, 0)
bench.apply_groups(ctx, groups)
queries, qgroups, _ = bench.select_queries(ctx, groups)
q = ctx.make_queries(queries, qgroups)
cache = os.path.join(ROOT, "scripts", "c4_pops_seed0.npy")
if os.path.exists(cache) and len(np.load(cache)) == len(q):
    pops = np.load(cache)
else:
    ctx.batch_upload(q, ctx.make_opts(path_cap=2048))
    ctx.batch_run()
    pops = ctx.batch_fetch()[0]["n_pops"]
    np.save(cache, pops)
keep = pops < a.max_pops
qs = q[keep]
ctx.batch_upload(qs, ctx.make_opts(path_cap=2048, max_slots=a.slots))
ms = ctx.batch_run()
tot = int(pops[keep].sum())
print(f"{len(qs)} queries, {tot} expansions, slots {a.slots}: {ms:.1f} ms = {tot / ms / 1e3:.3f} M exp/s")
