#!/usr/bin/env python
"""Kernel-level measurements for BASELINE configs C2 (map update at 2048^2, 256 boxes) and C3 (2D field + Dubins field
sweep over 2048^2 x 72), with roofline fractions against MEASURED_PEAKS.json.  One JSON line per kernel.
(The contract benchmark is bench.py; this script supplies the per-kernel numbers quoted in DESIGN.md / profiles/.)"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenarios as S  # noqa: E402
import path_planning_pkg_b200 as pp  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(ctx, fn, reps):
    ms = C.c_float()
    fn()
    ctx.sync()
    ctx.lib.pp_timer_begin(ctx.h)
    for _ in range(reps):
        fn()
    ctx.lib.pp_timer_end(ctx.h, C.byref(ms))
    return ms.value / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cpu", action="store_true", help="also time the reference (oracle/_ref) on one host core")
    a = ap.parse_args()
    peak, src = peaks()
    sc = S.c2_scenario(grid_size=a.n)
    P = pp.make_params(grid_size=a.n, resolution=sc["resolution"])
    ctx = pp.Context(P, num_groups=1)
    ctx.update_goal(sc["goal"], sc["frame_start"])
    nn = a.n * a.n
    out = []

    # C2: decay (pure streaming: 4 B read + 4 B write per cell)
    ms = timed(ctx, lambda: ctx.decay(), a.reps)
    gbs = 8.0 * nn / (ms * 1e-3) / 1e9
    out.append(dict(kernel="pp_map_decay_kernel", config=f"C2 decay {a.n}^2", ms=ms, algorithmic_bytes=8 * nn, achieved_gbs=gbs, peak_gbs=peak,
                    frac=gbs / peak, peak_source=src, note="16 MiB map is L2 resident on B200 (126 MB L2): can exceed the HBM roofline"))
    # C2: boxes (host prologue + H2D of descriptors + gather kernel + sync, as one call)
    ctx.set_map(np.zeros((a.n, a.n), np.float32))
    ms = timed(ctx, lambda: ctx.update_boxes_2d(sc["boxes"], sc["conf"]), max(a.reps // 4, 3))
    ctx.set_map(np.zeros((a.n, a.n), np.float32))
    ctx.update_boxes_2d(sc["boxes"], sc["conf"])
    touched = int((ctx.get_map() != 0).sum())
    alg = 8 * touched + 20 * len(sc["boxes"])
    out.append(dict(kernel="pp_map_update_kernel, boxes only (+host prologue, descriptor H2D; asynchronous)", config=f"C2 {len(sc['boxes'])} boxes into {a.n}^2", ms=ms,
                    distinct_cells=touched, algorithmic_bytes=alg, achieved_gbs=alg / (ms * 1e-3) / 1e9, peak_gbs=peak,
                    frac=alg / (ms * 1e-3) / 1e9 / peak, note="host-prologue bound: ~0.5 MB of useful traffic per call (SURVEY H4)"))
    # C2: one round = boxes + decay fused into a single pass over the map
    ms = timed(ctx, lambda: ctx.update_boxes_2d_decay(sc["boxes"], sc["conf"]), a.reps)
    alg = 8 * nn + 8 * touched + 32 * len(sc["boxes"])
    out.append(dict(kernel="pp_map_update_kernel, boxes + fused decay (one round, one launch)", config=f"C2 round {a.n}^2", ms=ms,
                    algorithmic_bytes=alg, achieved_gbs=alg / (ms * 1e-3) / 1e9, peak_gbs=peak, frac=alg / (ms * 1e-3) / 1e9 / peak,
                    note="per call from Python incl. the host prologue; the kernel alone is in the ncu launch list"))
    # three C2 rounds so the field kernels see the C2 map
    ctx.set_map(np.zeros((a.n, a.n), np.float32))
    for _ in range(sc["rounds"]):
        ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS)
        ctx.decay()
    # C3: 2D field
    f, sweeps, ms = ctx.field2d(download=True)
    f, sweeps, ms = ctx.field2d(download=False)
    alg = 8 * nn
    out.append(dict(kernel="pp_field2d_persistent_kernel (all sweeps, one launch)", config=f"C3 2D field {a.n}^2", ms=ms, sweeps=sweeps,
                    algorithmic_bytes_per_sweep=alg, achieved_gbs=alg * sweeps / (ms * 1e-3) / 1e9, peak_gbs=peak,
                    frac=alg * sweeps / (ms * 1e-3) / 1e9 / peak,
                    note="upper bound on traffic: inactive tiles are skipped, so real traffic per sweep is far below 8 B/cell"))
    # C3: Dubins field sweep (mandatory write: 4 B per state)
    _, ms = ctx.field3d(use_h2d=True, download=False)
    _, ms = ctx.field3d(use_h2d=True, download=False)
    states = nn * 72
    flops = 485.0 * states
    out.append(dict(kernel="pp_dubins_field_kernel", config=f"C3 Dubins sweep {a.n}^2 x 72", ms=ms, states=states,
                    states_per_s=states / (ms * 1e-3), write_gbs=4.0 * states / (ms * 1e-3) / 1e9, write_frac_of_hbm=4.0 * states / (ms * 1e-3) / 1e9 / peak,
                    fp32_tflops_est=flops / (ms * 1e-3) / 1e12, fp32_peak_tflops_derived=74.4, fp32_frac=flops / (ms * 1e-3) / 1e12 / 74.4,
                    note="485 FP32 op/state is SURVEY 8d's estimate (125 basic + 18 transcendentals x 20); FP32 peak derived, not measured"))
    # FP32 SIMT Dubins length per state (K-POP heuristic flavour), through the C ABI (H2D 12 B + D2H 4 B per state included)
    m = 1 << 22
    starts = np.random.RandomState(0).uniform(0, a.n * sc["resolution"], (m, 3)).astype(np.float32)
    goal = np.array(list(ctx.frame().goal_grid), np.float32)
    ms = timed(ctx, lambda: ctx.dubins_length_fp32(starts, goal), 3)
    out.append(dict(kernel="pp_dubins_length_fp32_kernel (+H2D 12 B, D2H 4 B per state, pageable host memory)", config=f"{m} states",
                    ms=ms, states_per_s=m / (ms * 1e-3)))
    # north_star (c): generic footprint collision kernel on this map (4.0 x 2.0 m rectangle), kernel time alone
    try:
        mfp = 1 << 20
        L = a.n * sc["resolution"]
        rs = np.random.RandomState(3)
        fp = np.concatenate([rs.uniform(0.0, L, (mfp, 2)), rs.uniform(-np.pi, np.pi, (mfp, 1))], 1).astype(np.float32)
        ctx.set_footprint(4.0, 2.0, 1.0)
        ctx.footprint(fp)
        fms = min(ctx.footprint(fp, want_ms=True)[3] for _ in range(5))
        cells = float(np.mean([len(ctx.footprint_table(b)) for b in range(72)]))
        out.append(dict(kernel="pp_footprint_kernel", config=f"{mfp} poses, 4x2 m rectangle on {a.n}^2", ms=fms, poses_per_s=mfp / (fms * 1e-3),
                        mean_cells_per_pose=cells, algorithmic_gbs=mfp * (cells * 4 + 24) / (fms * 1e-3) / 1e9, peak_gbs=peak,
                        note="map reads are served by L1/L2 (DESIGN.md section 12)"))
    except Exception as e:     # a side measurement of this script
        out.append(dict(kernel="pp_footprint_kernel", error=str(e)))
    # stateless velocity-profile kernel (one thread per path, 64-point paths), host buffers included
    try:
        npth, cap = 1 << 16, 64
        rs = np.random.RandomState(4)
        xy = np.cumsum(rs.uniform(0.1, 0.4, (npth, cap, 2)), 1).astype(np.float32)[:, ::-1].copy()
        cv = np.abs(rs.uniform(-0.2, 0.2, (npth, cap))).astype(np.float32)
        cnt = np.full(npth, cap, np.int32); vi = rs.uniform(0, 2, npth).astype(np.float32)
        lim = [5.0, 1.0, 2.0, 1.0, 2.5]
        ctx.velocity_profile_batch(lim, xy, cv, cnt, vi)
        t = time.perf_counter(); ctx.velocity_profile_batch(lim, xy, cv, cnt, vi); vms = (time.perf_counter() - t) * 1e3
        out.append(dict(kernel="pp_velocity_profile_kernel (+H2D / D2H, pageable host memory)", config=f"{npth} paths x {cap} points", ms=vms,
                        paths_per_s=npth / (vms * 1e-3)))
    except Exception as e:
        out.append(dict(kernel="pp_velocity_profile_kernel", error=str(e)))
    if a.cpu:
        import orc
        o = orc.ref(orc.make_params(grid_size=a.n, resolution=sc["resolution"])) if orc.have_ref() else orc.port(orc.make_params(grid_size=a.n, resolution=sc["resolution"]))
        o.update_goal(sc["goal"], sc["frame_start"])
        t = time.perf_counter(); o.update_boxes_2d(sc["boxes"], sc["conf"]); tb = time.perf_counter() - t
        t = time.perf_counter(); o.decay(); td = time.perf_counter() - t
        goal = np.array(list(o.consts().goal_grid), np.float32)
        m = 200000
        starts = np.random.RandomState(0).uniform(0, a.n * sc["resolution"], (m, 3)).astype(np.float32)
        t = time.perf_counter(); o.dubins_length(starts, goal); tdu = (time.perf_counter() - t) / m
        out.append(dict(kernel="cpu reference, 1 core", boxes_ms=tb * 1e3, decay_ms=td * 1e3, dubins_ns_per_state=tdu * 1e9,
                        dubins_sweep_extrapolated_s=tdu * states))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
