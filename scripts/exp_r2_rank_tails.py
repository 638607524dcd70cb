#!/usr/bin/env python
"""Development probe: the C4 batches of ranks 0..7 (seeds r*64 .. r*64+63) one after the other on one GPU: expansions per batch,
the longest queries, isolated batch time.  Tells how long the slowest rank's drain is at N = 8 (bench.py's per-rank workloads)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import path_planning_pkg_b200 as pp  # noqa: E402
import bench  # noqa: E402

P = pp.make_params(grid_size=512, resolution=0.2)
for r in range(8):
    ctx = pp.Context(P, num_groups=64)
    groups = bench.build_workload(64, 64, r * 64)
    bench.apply_groups(ctx, groups)
    queries, qgroups, _ = bench.select_queries(ctx, groups)
    q = ctx.make_queries(queries, qgroups)
    ctx.batch_upload(q, ctx.make_opts(path_cap=2048))
    ms = ctx.batch_run()
    res, _, _ = ctx.batch_fetch()
    n = np.sort(res["n_pops"])[::-1]
    print(f"rank {r}: {len(q)} queries, {int(n.sum())} expansions, batch alone {ms / 1e3:.1f} s, longest queries {n[:4].tolist()}, "
          f"> 200k: {int((n > 200000).sum())}, retried {ctx.batch_retried()}, status flags {int((res['status'] != 0).sum())}", flush=True)
    ctx.close()
