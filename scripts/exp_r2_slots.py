#!/usr/bin/env python
"""Development experiment (round 2): bulk throughput of the EXACT search kernel against the number of resident queries.
The C4 batch without its long tail (queries above --max-pops expansions dropped, so that the launch time is the bulk, not the one
longest query), one lane, slots swept.  Prints expansions/s per configuration."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import path_planning_pkg_b200 as pp  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=64)
    ap.add_argument("--max-pops", type=int, default=60000)
    ap.add_argument("--slots", default="148,296,592,888,1184,1776,2368")
    a = ap.parse_args()
    P = pp.make_params(grid_size=512, resolution=0.2)
    ctx = pp.Context(P, num_groups=a.groups)
    groups = bench.build_workload(a.groups, 64, 0)
    bench.apply_groups(ctx, groups)
    queries, qgroups, _ = bench.select_queries(ctx, groups)
    q = ctx.make_queries(queries, qgroups)
    ctx.batch_upload(q, ctx.make_opts(path_cap=2048))
    ms = ctx.batch_run()
    res, _, _ = ctx.batch_fetch()
    print(f"full batch: {len(q)} queries, {int(res['n_pops'].sum())} expansions, {ms:.0f} ms", flush=True)
    keep = res["n_pops"] < a.max_pops
    qs = q[keep]
    pops = int(res["n_pops"][keep].sum())
    # longest first within the kept set would hide imbalance; keep the natural order and make the batch big enough instead
    print(f"kept {len(qs)} queries below {a.max_pops} expansions: {pops} expansions (max {int(res['n_pops'][keep].max())})", flush=True)
    for s in [int(v) for v in a.slots.split(",")]:
        ctx.batch_upload(qs, ctx.make_opts(path_cap=2048, max_slots=s))
        ms = min(ctx.batch_run() for _ in range(2))
        print(f"slots {s:5d}: {ms:9.1f} ms  {pops / ms / 1e3:8.3f} M exp/s  per-warp {pops / ms * 1e3 / s / 1e3:7.2f} k exp/s "
              f"({1e6 * s * ms * 1e-3 / pops:6.1f} us/expansion/warp)", flush=True)


if __name__ == "__main__":
    main()
