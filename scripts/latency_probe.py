#!/usr/bin/env python
"""Development probe: single-query latency through the C ABI on the C1 scenarios, EXACT and K-POP(32) modes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenarios as S  # noqa: E402
import path_planning_pkg_b200 as pp  # noqa: E402


def main():
    seeds = list(range(int(sys.argv[1]) if len(sys.argv) > 1 else 40))
    scs = [S.c1_scenario(s) for s in seeds]
    P = pp.make_params(grid_size=scs[0]["grid_size"], resolution=scs[0]["resolution"])
    ctx = pp.Context(P, num_groups=len(seeds))
    qs = []
    for gi, sc in enumerate(scs):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        for _ in range(sc["rounds"]):
            ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
            ctx.decay(group=gi)
        qs.append(sc["queries"][0])
    q = ctx.make_queries(np.array(qs), list(range(len(seeds))))
    for mode, name in ((0, "EXACT"), (1, "K-POP(32)")):
        for cap in (1 << 17, 1 << 13):
            opts = ctx.make_opts(path_cap=1024, max_slots=1, mode=mode, kpop=32, max_expansions=cap, max_open=cap // 2)
            ctx.find_path_batch(q[:1], opts)          # pools, field
            lat, pops, kms = [], [], []
            for k in range(len(q)):
                t0 = time.perf_counter()
                res, _, _, _ = ctx.find_path_batch(q[k:k + 1], opts)
                lat.append((time.perf_counter() - t0) * 1e3); pops.append(int(res[0]["n_pops"]))
                ctx.batch_upload(q[k:k + 1], opts); kms.append(ctx.batch_run())
            lat, pops, kms = np.array(lat), np.array(pops), np.array(kms)
            print(f"{name:10s} cap {cap:7d}: call p50 {np.median(lat):7.3f} ms p95 {np.percentile(lat, 95):7.3f} | kernel p50 {np.median(kms):7.3f} ms | "
                  f"pops p50 {int(np.median(pops))} | min call {lat.min():.3f} ms (pops {pops[lat.argmin()]})")


if __name__ == "__main__":
    main()
