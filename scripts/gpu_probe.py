#!/usr/bin/env python
"""Development probe (not part of the product or the tests): runs a C4-shaped batch and prints work statistics
and, when PP_B200_LIB points at the -DPP_PROFILE variant, the per-phase cycle split of the search kernel."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenarios as S  # noqa: E402
import path_planning_pkg_b200 as pp  # noqa: E402
import bench  # noqa: E402

PHASES = ["init", "pop+closed.insert+erase", "rollout+collision+apf", "dubins cand", "closed.find", "spec walks + find commit",
          "lazy 2D A*", "insert commit"]
KPHASES = ["init", "queue pop (select k)", "validate+close", "expand: rollout+collision+apf+hash", "winners: dubins+node write",
           "sort + LSM insert", "goal / dubins shot", "-"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, default=8)
    ap.add_argument("--starts", type=int, default=64)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--max-expansions", type=int, default=0)
    ap.add_argument("--max-open", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--dump", default="", help="write per-query statistics to this .npz")
    a = ap.parse_args()
    P = pp.make_params(grid_size=512, resolution=0.2)
    ctx = pp.Context(P, num_groups=a.groups)
    groups = bench.build_workload(a.groups, a.starts, 0)
    bench.apply_groups(ctx, groups)
    queries, qgroups, maps = bench.select_queries(ctx, groups)
    q = ctx.make_queries(queries, qgroups)
    opts = ctx.make_opts(max_expansions=a.max_expansions, max_open=a.max_open, max_slots=a.slots, mode=a.mode, kpop=a.k)
    ctx.batch_upload(q, opts)
    prof = hasattr(ctx.lib, "pp_profile_read")
    buf = (C.c_ulonglong * 16)()
    if prof:
        ctx.lib.pp_profile_read(ctx.h, buf)
    for rep in range(a.reps):
        ms = ctx.batch_run()
        res, _, _ = ctx.batch_fetch()
        pops = int(res["n_pops"].sum())
        print(f"rep {rep}: {len(q)} queries, {pops} expansions, {ms:.1f} ms -> {pops / ms / 1e3:.3f} M exp/s, "
              f"{len(q) / ms * 1e3:.1f} q/s")
    np_ = res["n_pops"]
    print("pops/query: min %d p50 %d p90 %d max %d mean %.0f" % (np_.min(), np.median(np_), np.percentile(np_, 90), np_.max(), np_.mean()))
    print("max_open: p50 %d max %d | n_closed max %d | lazy searches/query mean %.0f | lazy pops/3D pop %.2f" % (
        np.median(res["max_open"]), res["max_open"].max(), res["n_closed"].max(), res["n_lazy_searches"].mean(),
        res["n_lazy_pops"].sum() / max(pops, 1)))
    print("success %.3f, status!=0: %d, oob pops %d" % (res["success"].mean(), (res["status"] != 0).sum(), res["n_pops_bin_oob"].sum()))
    if a.mode == 1:
        it = res["n_lazy_searches"].astype(np.float64); tk = res["n_lazy_pops"].astype(np.float64)
        print("kpop: iterations/query mean %.0f max %d | valid pops per iteration %.2f | entries taken per iteration %.2f | nodes max %d" % (
            it.mean(), it.max(), np_.sum() / it.sum(), tk.sum() / it.sum(), res["n_closed"].max()))
        order = np.argsort(-np_)[:5]
        print("longest queries: pops", np_[order], "iterations", res["n_lazy_searches"][order], "success", res["success"][order])
    if a.dump:
        h1 = np.zeros(len(q), np.float32)
        st = ctx.set_start(q)
        for g in range(a.groups):
            f, _, _ = ctx.field2d(group=g)
            m = qgroups == g
            h1[m] = f[st["ci"][m], st["cj"][m]]
        np.savez_compressed(a.dump, queries=queries, qgroups=qgroups, n_pops=res["n_pops"], iters=res["n_lazy_searches"], success=res["success"],
                            cost=res["cost"], h1=h1, n_nodes=res["n_closed"])
    if prof and a.mode == 1:
        t0 = res["max_open"].astype(np.int64); t1 = res["n_pops_bin_oob"].astype(np.int64)
        base = t0.min(); t0 -= base; t1 -= base
        dur = np.maximum(t1 - t0, 1); its = res["n_lazy_searches"].astype(np.float64)
        print("timeline (us): last start %d, last end %d; queries ending in the last 10%% of the batch: %d" % (
            t0.max(), t1.max(), (t1 > 0.9 * t1.max()).sum()))
        edges = np.linspace(0, t1.max(), 11)
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = (t0 < hi) & (t1 > lo)                      # queries alive in this window
            sel = (t0 >= lo) & (t0 < hi) & (its > 50)
            rate = (dur[sel] / its[sel]).mean() if sel.any() else float("nan")
            print("  window %7.0f-%7.0f us: alive %4d, started %4d, mean us/iteration of those started here %.1f" % (lo, hi, m.sum(), ((t0 >= lo) & (t0 < hi)).sum(), rate))
        order = np.argsort(-t1)[:5]
        print("  last finishers: start", t0[order], "end", t1[order], "iterations", res["n_lazy_searches"][order])
    if prof:
        ctx.lib.pp_profile_read(ctx.h, buf)
        tot = sum(buf[:8])
        for name, v in zip(KPHASES if a.mode == 1 else PHASES, buf[:8]):
            print(f"  {name:28s} {100.0 * v / max(tot, 1):6.2f} %   {v / max(pops, 1):10.0f} cycles/expansion")
        if a.mode == 0:
            names = ["find walks used as speculated", "find walks redone", "insert walks speculated + used", "insert walks speculated, redone",
                     "insert walks without speculation", "successors committed", "inserts attached", "mutation-log entries seen"]
            for name, v in zip(names, buf[8:]):
                print(f"  {name:36s} {v / max(pops, 1):8.3f} per expansion")


if __name__ == "__main__":
    main()
