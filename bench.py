#!/usr/bin/env python
"""bench.py -- batched Hybrid A* queries on synthetic clutter maps (BASELINE.json configs[3], with configs[0], [1], [2], [4] as blocks).

Headline (`value`, `e2e`): one "step" = one pass of the hot path over one batch = 64 (map, goal) groups x 64 start poses = 4096
independent HybridAStar::find_path queries on 512 x 512 x 72 maps with 96 box obstacles each, EXACT single-pop
mode: every query returns what the unmodified reference returns (expansion sequence, cost, path).  Steps are independent batches,
so they are submitted the way a serving system would: round-robin over `--lanes` lane contexts (pp_create_lane: own stream and
scratch, shared maps), one synchronisation at the end -- the drain of batch k (its few longest queries, one warp each) overlaps
the bulk of batch k+1.  `batch_latency_ms` is the same batch alone on an idle GPU.

  value : node expansions / s, whole job, queries resident in HBM, K steps bracketed by CUDA events on the context's stream
  e2e   : the same through the C ABI with pinned HOST buffers every step (H2D queries, D2H results + paths + curvature)
  --gpus N (torchrun): weak scaling, rank r runs groups with seeds r*64 .. r*64+63 (distinct work), no data-path collective
  c5 block : BASELINE configs[4]: 65 536 queries = 1 024 groups x 64 starts IN TOTAL, K-POP(32), sharded over the ranks (strong
             scaling); rank 0 rasterises every map, pp_broadcast_maps (NCCL) replicates them into every rank's context
  c1 / c2_map_update / c3_fields blocks : configs[0], [1], [2] (single-query latency, map update round, heuristic field sweep)
  --impl reference : the unmodified reference (oracle/_ref, stock libm) on all host cores, bounded sample per step
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import scenarios as S  # noqa: E402  (generators only, no oracle code)

N_GRID, RES = 512, 0.2
ALGO_BYTES_PER_EXPANSION = 312   # state 32 B + 5 x (map 4 + closed probe 16 + open insert 32 + h 4)
T_START = time.time()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def elapsed():
    return time.time() - T_START


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(n_groups, n_starts, seed0):
    return [S.c4_group(seed0 + g, n_starts=n_starts, grid_size=N_GRID, resolution=RES) for g in range(n_groups)]


def apply_groups(ctx, groups, rasterise=True):
    """update_goal + 4 x (boxes, decay) per group through the C ABI (rasterise=False: frames and APF lists only, the maps arrive
    by broadcast)."""
    for gi, sc in enumerate(groups):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        if rasterise:
            for _ in range(sc["rounds"]):
                ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
                ctx.decay(group=gi)
        else:
            ctx.update_apf(sc["boxes"], S.APF_ADDED_RADIUS, group=gi)


def select_queries(ctx, groups, maps=None):
    """Start poses in free cells (reject starts in occupied cells), per group, from the device's own maps."""
    thr = ctx.consts().log_threshold
    queries, qgroups, out_maps = [], [], []
    for gi, sc in enumerate(groups):
        m = ctx.get_map(gi)
        out_maps.append(m)
        cand = sc["start_candidates"]
        st = ctx.set_start(ctx.make_queries(cand, [gi] * len(cand)))
        free = m[st["ci"], st["cj"]] < thr
        sel = cand[free][:sc["n_starts"]]
        queries.append(sel); qgroups += [gi] * len(sel)
    return np.concatenate(queries), np.array(qgroups, np.int32), out_maps


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def committed_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r2_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[kernel]
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
class Pipeline:
    """K independent batches over S lane contexts: submit round-robin, collect a lane's previous batch before reusing it."""

    def __init__(self, ctx, lanes, q, opts, pp, fetch_lanes=None):
        """fetch_lanes: how many lanes (the first ones) ever run end-to-end steps; only those get pinned host buffers for the results
        (134 MB per lane at C4 size -- 20 lanes x 8 ranks would pin 21 GB for buffers the device-resident steps never touch)."""
        import torch
        self.pp, self.ctx, self.q, self.opts, self.n = pp, ctx, q, opts, len(q)
        self.lanes = [ctx] + [ctx.create_lane() for _ in range(max(lanes, 1) - 1)]
        self.busy = [False] * len(self.lanes)
        self.kernel_ms = []
        pc = opts.path_cap
        n_f = len(self.lanes) if fetch_lanes is None else max(1, min(len(self.lanes), int(fetch_lanes)))
        self.hq = torch.from_numpy(q.view(np.uint8).copy()).pin_memory()
        self.hres = [torch.zeros(self.n * pp._cabi.RESULT_DT.itemsize, dtype=torch.uint8).pin_memory() for _ in range(n_f)]
        self.hpath = [torch.zeros(self.n * pc * 3, dtype=torch.float32).pin_memory() for _ in range(n_f)]
        self.hcurv = [torch.zeros(self.n * pc, dtype=torch.float32).pin_memory() for _ in range(n_f)]
        self.h2d = int(q.nbytes)
        self.d2h = int(self.hres[0].numel() + self.hpath[0].numel() * 4 + self.hcurv[0].numel() * 4)

    def set_budget(self, nbytes):
        for l in self.lanes:
            l.set_memory_budget(nbytes)

    def upload_all(self):
        for l in self.lanes:
            l.batch_upload(self.q, self.opts)

    def _collect(self, i, e2e):
        l = self.lanes[i]
        self.kernel_ms.append(l.batch_wait())
        if e2e:
            if i >= len(self.hres):
                raise RuntimeError(f"lane {i} has no host buffers (fetch_lanes = {len(self.hres)})")
            rc = l.lib.pp_batch_fetch(l.h, C.c_void_p(self.hres[i].data_ptr()), C.c_void_p(self.hpath[i].data_ptr()),
                                      C.c_void_p(self.hcurv[i].data_ptr()), None)
            if rc != 0:
                raise RuntimeError(l.lib.pp_last_error().decode())
        self.busy[i] = False

    def run(self, steps, e2e, mark_at=0):
        """Submits `steps` batches round-robin and collects them all; returns the host time at which step `mark_at` was submitted
        (the start of a timed region that begins with the pipeline already full)."""
        self.kernel_ms = []
        t_mark = None
        for k in range(steps):
            if k == mark_at:
                t_mark = time.perf_counter()
            i = k % len(self.lanes)
            if self.busy[i]:
                self._collect(i, e2e)
            l = self.lanes[i]
            if e2e:
                rc = l.lib.pp_batch_upload(l.h, C.c_void_p(self.hq.data_ptr()), C.c_int(self.n), C.byref(self.opts))
                if rc != 0:
                    raise RuntimeError(l.lib.pp_last_error().decode())
                l._opts, l._n = self.opts, self.n
            l.batch_run_async()
            self.busy[i] = True
        for i in range(len(self.lanes)):
            if self.busy[i]:
                self._collect(i, e2e)
        return t_mark

    def results(self, lane=0):
        return np.frombuffer(self.hres[lane].numpy().tobytes(), self.pp._cabi.RESULT_DT)

    def close(self):
        for l in self.lanes[1:]:
            l.close()
        self.lanes = self.lanes[:1]


def timed(ctx, barrier, fn):
    """fn() bracketed by barrier + synchronize and by CUDA events on the context's stream; returns (event ms, wall s)."""
    barrier()
    ctx._chk(ctx.lib.pp_timer_begin(ctx.h))
    t0 = time.perf_counter()
    fn()
    ms = C.c_float()
    ctx._chk(ctx.lib.pp_timer_end(ctx.h, C.byref(ms)))
    wall = time.perf_counter() - t0
    barrier()
    return float(ms.value), wall


# ---------------------------------------------------------------------------------------------------------------------
def reference_sample(nq, n, seed=1):
    """The bounded CPU sample both CPU legs draw from: a fixed permutation of the batch's query indices."""
    return np.random.RandomState(seed).permutation(nq)[:n]


def run_reference_arm(args, workload):
    """--impl reference: the unmodified reference on every host core, one planner per thread, scrubbed per query (F12)."""
    import orc
    n_threads = os.cpu_count() or 1
    groups = build_workload(args.groups, args.starts, 0)
    P = orc.make_params(grid_size=N_GRID, resolution=RES)
    maps, queries, qgroups = [], [], []
    o = orc.ref(P)
    for gi, sc in enumerate(groups):      # maps and start selection from the reference itself (no GPU on this arm)
        o.set_map(np.zeros((N_GRID, N_GRID), np.float32))
        S.build_map(o, sc)
        m = o.get_map(); maps.append(m)
        sel = S.select_starts(sc, m, o.consts().log_threshold, o.set_start)
        queries.append(sel); qgroups += [gi] * len(sel)
    queries = np.concatenate(queries); qgroups = np.array(qgroups, np.int32)
    per_step = int(min(max(8 * n_threads, 64), 256, len(queries)))
    order = reference_sample(len(queries), len(queries))
    tot_pops, tot_s, tot_busy, tot_q, k = 0, 0.0, 0.0, 0, 0
    for step in range(args.warmup + args.steps):
        if step < args.warmup:          # warm-up: page in the library, touch the planners' memory; a handful of queries is enough
            idx = order[:max(n_threads, 8)]
        else:
            idx = order[(k * per_step) % len(order):][:per_step]; k += 1
        b = orc.ref_batch(P, groups, queries, qgroups, maps, idx, n_threads)
        if step >= args.warmup:
            tot_pops += int(b["pops"].sum()); tot_s += b["secs"]; tot_busy += float(b["busy_s"].sum()); tot_q += len(idx)
    val = tot_pops / tot_s
    line = {"impl": "reference", "metric": "hybrid_astar_node_expansions_per_s", "value": val, "unit": "expansions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "queries_per_s": tot_q / tot_s,
            "busy_time": {"expansions_per_s_per_core": tot_pops / tot_busy, "expansions_per_s_all_cores_no_idle": tot_pops / tot_busy * n_threads,
                          "note": "expansions / time spent inside find_path summed over threads: what the cores deliver when none waits for "
                                  "the step's longest query"},
            "config": {"workload": workload, "sample": f"{per_step} queries per step in the fixed permutation (seed 1) the GPU arm's cpu_baseline uses"},
            "cpu_baseline": {"value": val, "unit": "expansions/s", "cores": n_threads, "kind": "reference",
                             "sample": f"{per_step} queries/step x {args.steps} steps, one HybridAStar<float> per thread, scrubbed per query"},
            "e2e": {"value": val, "unit": "expansions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def block_c5(args, pp, torch, dist, rank, local_rank, world, barrier, in_time):
    """BASELINE configs[4] / SURVEY 8d C5: 65 536 queries = 1 024 groups x 64 starts in total, K-POP(32), sharded by query over the
    ranks (strong scaling).  Rank 0 rasterises every map; pp_broadcast_maps replicates them INTO every rank's context (NCCL)."""
    G = args.c5_groups
    P = pp.make_params(grid_size=N_GRID, resolution=RES)
    ctx = pp.Context(P, num_groups=G, device=local_rank)
    groups = [S.c4_group(g, n_starts=args.starts, grid_size=N_GRID, resolution=RES) for g in range(G)]
    t0 = time.time()
    apply_groups(ctx, groups, rasterise=(rank == 0))
    ctx.sync()
    build_s = time.time() - t0
    bcast_ms = None
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        ctx.broadcast_maps(0, G, 0); ctx.sync()                 # communicator warm-up (and the replication itself)
        ms, _ = timed(ctx, barrier, lambda: (ctx.broadcast_maps(0, G, 0), ctx.sync()))
        bcast_ms = ms
        # replicas must be bit-identical to what this rank would have rasterised itself: spot-check one group per rank
        g = (rank * 7 + 3) % G
        probe = pp.Context(P, num_groups=1, device=local_rank)
        apply_groups(probe, [groups[g]], rasterise=True)
        same = bool(np.array_equal(probe.get_map(0).view(np.uint32), ctx.get_map(g).view(np.uint32)))
        probe.close()
        flag = torch.tensor([1.0 if same else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        replicas_ok = bool(flag.item() == 1.0)
    else:
        replicas_ok = None
    queries, qgroups, _ = select_queries(ctx, groups)
    mine = np.arange(rank, len(queries), world)       # every world-th query (path_planning_pkg_b200.shard.shard_queries): equal work per rank
    total_q = len(queries)
    q = ctx.make_queries(queries[mine], qgroups[mine])
    opts = ctx.make_opts(max_expansions=1 << 18, path_cap=1024, mode=1, kpop=32)
    ctx.batch_upload(q, opts)
    ctx.batch_run()                                            # warm-up (also builds the scheduling hints)
    steps = args.c5_steps if in_time() else 1
    l0 = ctx.kernel_launches()
    kms = []
    ms, wall = timed(ctx, barrier, lambda: kms.extend(ctx.batch_run() for _ in range(steps)))
    launches = ctx.kernel_launches() - l0
    res, _, _ = ctx.batch_fetch()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(res["n_pops"].sum()), float(len(q)), float(res["success"].sum()), float((res["status"] != 0).sum())],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ctx.close()
    max_ms = float(t.item()); pops, nq, succ, flags = [float(v) for v in cnt.tolist()]
    return {"workload": f"C5: {G} groups x {args.starts} starts = {int(nq)} queries in total over {world} GPU(s), {N_GRID}x{N_GRID}x72, K-POP(32)",
            "scaling": "strong", "steps": steps, "ms_per_step": max_ms / steps, "queries_per_s": nq * steps / (max_ms * 1e-3),
            "expansions_per_s": pops * steps / (max_ms * 1e-3), "queries_total": int(total_q), "success_rate": succ / max(nq, 1),
            "capacity_flags": int(flags), "gpu_launches": int(launches),
            "map_replication": {"how": "rank 0 rasterises all maps; pp_broadcast_maps (ncclBroadcast) writes them into every rank's context",
                                "bytes": int(G) * N_GRID * N_GRID * 4, "broadcast_ms": bcast_ms, "replicas_bit_identical": replicas_ok,
                                "map_build_s_rank0": build_s},
            "note": "K-POP has its own semantics (DESIGN.md section 9): bit-identical to its CPU restatement, not to the reference"}


def block_kpop(args, pp, ctx, q, res_exact, barrier, cpu, have_time):
    """K-POP(32) on the headline batch + its cost deviation from the REFERENCE on the CPU sample (not from the EXACT mode)."""
    kopts = ctx.make_opts(max_expansions=1 << 18, path_cap=2048, mode=1, kpop=32)
    ctx.batch_upload(q, kopts)
    ctx.batch_run()
    steps = 3 if have_time else 1
    kms = []
    ms, _ = timed(ctx, barrier, lambda: kms.extend(ctx.batch_run() for _ in range(steps)))
    kres, _, _ = ctx.batch_fetch()
    out = {"k": 32, "steps": steps, "ms_per_step": ms / steps, "expansions_per_s": float(kres["n_pops"].sum()) * steps / (ms * 1e-3),
           "queries_per_s": len(q) * steps / (ms * 1e-3), "success_rate": float(kres["success"].mean()),
           "note": "k pops per iteration, exact 2D field heuristic, no equal-f drops: own semantics, bit-identical to its CPU restatement "
                   "(oracle/port/kpop.inc); north_star's 1e-4 cost bar against the reference is NOT met and cannot be (SURVEY F4/F5)"}
    if cpu is not None:
        idx, b = cpu
        both = (b["success"] == 1) & (kres["success"][idx] == 1)
        dev = kres["cost"][idx][both] / b["cost"][both] - 1.0
        out["cost_vs_reference"] = {"sample": f"{int(both.sum())} queries of the cpu_baseline sample both solved",
                                    "median": float(np.median(dev)), "min": float(dev.min()), "max": float(dev.max()),
                                    "within_1e-4": int((np.abs(dev) <= 1e-4).sum()),
                                    "success_agree": int((b["success"] == kres["success"][idx]).sum())}
    return out


def block_c1(args, pp, local_rank, n_seeds=32):
    """BASELINE configs[0] / SURVEY 8d C1: single local_planner query, N=200 @0.2 m, 5 boxes; GPU latency through the C ABI next to
    the reference on one host core, seeds 0..n-1, plus identity of the results."""
    import orc
    scs = [S.c1_scenario(s) for s in range(n_seeds)]
    P = orc.make_params(grid_size=scs[0]["grid_size"], resolution=scs[0]["resolution"])
    ctx = pp.Context(pp._cabi.params_from(P), num_groups=n_seeds, device=local_rank)
    ref = orc.ref(P)
    for gi, sc in enumerate(scs):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        for _ in range(sc["rounds"]):
            ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
            ctx.decay(group=gi)
    q = ctx.make_queries(np.array([sc["queries"][0] for sc in scs]), list(range(n_seeds)))
    o1 = ctx.make_opts(path_cap=2048, max_slots=1)
    ctx.find_path_batch(q[:1], o1)
    gpu_ms, cpu_ms, same, pops = [], [], 0, []
    for gi, sc in enumerate(scs):
        t0 = time.perf_counter(); r, paths, curv, _ = ctx.find_path_batch(q[gi:gi + 1], o1); gpu_ms.append((time.perf_counter() - t0) * 1e3)
        ref.set_map(np.zeros((sc["grid_size"], sc["grid_size"]), np.float32)); S.build_map(ref, sc); ref.scrub()
        qq = sc["queries"][0]
        t0 = time.perf_counter(); b = ref.find_path(float(qq[3]), qq[:3]); cpu_ms.append((time.perf_counter() - t0) * 1e3)
        n = int(r[0]["n_path"])
        same += int(int(r[0]["n_pops"]) == b["n_pops"] and np.float32(r[0]["cost"]) == b["cost"] and
                    np.array_equal(paths[0, :n].view(np.uint32), b["path"].view(np.uint32)))
        pops.append(int(r[0]["n_pops"]))
    ctx.close()
    return {"workload": f"C1: N=200, res 0.2, 72 bins, 5 boxes, 4 rounds, seeds 0-{n_seeds - 1}, one find_path at a time, EXACT mode",
            "gpu_ms": {"p50": float(np.median(gpu_ms)), "p95": float(np.percentile(gpu_ms, 95))},
            "reference_1core_ms": {"p50": float(np.median(cpu_ms)), "p95": float(np.percentile(cpu_ms, 95))},
            "expansions_p50": float(np.median(pops)), "identical_to_reference": f"{same}/{n_seeds}"}


def block_c2(args, pp, local_rank, peak):
    """BASELINE configs[1] / SURVEY 8d C2: 256 boxes into a 2048^2 log-odds map + decay = one round; device time per round."""
    sc = S.c2_scenario()
    P = pp.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx = pp.Context(P, num_groups=1, device=local_rank)
    ctx.update_goal(sc["goal"], sc["frame_start"])
    ctx.update_boxes_2d_decay(sc["boxes"], sc["conf"]); ctx.sync()
    reps = 50
    ms = C.c_float()
    ctx._chk(ctx.lib.pp_timer_begin(ctx.h))
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.update_boxes_2d_decay(sc["boxes"], sc["conf"])
    ctx._chk(ctx.lib.pp_timer_end(ctx.h, C.byref(ms)))
    host_ms = (time.perf_counter() - t0) * 1e3 / reps
    per = ms.value / reps
    # the two-call form the reference's caller uses (src/local_planner.cpp:241 then :288)
    ms2 = C.c_float()
    ctx._chk(ctx.lib.pp_timer_begin(ctx.h))
    for _ in range(reps):
        ctx.update_boxes_2d(sc["boxes"], sc["conf"]); ctx.decay()
    ctx._chk(ctx.lib.pp_timer_end(ctx.h, C.byref(ms2)))
    per2 = ms2.value / reps
    nn = sc["grid_size"] ** 2
    algo = 2 * 4 * nn + 470488 + 20 * len(sc["boxes"])          # fused single pass = 33.56 MB
    ctx.close()
    return {"workload": "C2: 256 boxes rasterised into a 2048x2048 log-odds map + whole-map decay (one round)", "round_ms": per,
            "algorithmic_bytes": algo, "achieved_GBps": algo / (per * 1e-3) / 1e9, "frac_of_hbm_peak": algo / (per * 1e-3) / 1e9 / peak,
            "host_ms_per_call": host_ms, "two_call_round_ms": per2,
            "note": "round = pp_update_obstacles_boxes_2d_decay (one fused pass, device-side binning, no synchronisation) through the C ABI with "
                    "host boxes; device time per round over 50 back-to-back rounds incl. the descriptor H2D; the 16 MiB map is L2-resident"}


def block_c3(args, pp, local_rank, peak):
    """BASELINE configs[2] / SURVEY 8d C3: exact 2D field + Dubins field over 2048 x 2048 x 72 for one goal."""
    sc = S.c2_scenario()
    P = pp.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx = pp.Context(P, num_groups=1, device=local_rank)
    ctx.update_goal(sc["goal"], sc["frame_start"])
    for _ in range(3):
        ctx.update_boxes_2d(sc["boxes"], sc["conf"]); ctx.decay()
    ctx.field2d(download=False)
    f2 = [ctx.field2d(download=False) for _ in range(3)]
    ctx.field3d(download=False)
    f3 = [ctx.field3d(download=False)[1] for _ in range(3)]
    ctx.close()
    states = sc["grid_size"] ** 2 * 72
    ms3 = float(np.min(f3))
    return {"workload": "C3: 2D holonomic field + Dubins field sweep for one goal over 2048x2048x72",
            "field2d_ms": float(np.min([f[2] for f in f2])), "field2d_sweeps": int(f2[0][1]),
            "dubins_field_ms": ms3, "states_per_s": states / (ms3 * 1e-3),
            "dubins_write_GBps": states * 4 / (ms3 * 1e-3) / 1e9, "dubins_write_frac_of_hbm_peak": states * 4 / (ms3 * 1e-3) / 1e9 / peak,
            "fp32_tflops_at_485_op_per_state": states * 485 / (ms3 * 1e-3) / 1e12}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--groups", type=int, default=64)
    ap.add_argument("--starts", type=int, default=64)
    ap.add_argument("--lanes", type=int, default=20, help="lane contexts the steps are submitted over (1 = one batch at a time), never more "
                                                          "than the timed steps.  With every timed batch in flight each batch's longest "
                                                          "queries start at once and run beside the bulk of all the others; with fewer lanes "
                                                          "a lane's next batch waits for the drain of its previous one")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0 = 8 per host thread, 64..256)")
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (0 = lanes / 2 at N = 1, lanes / 4 at N > 1 where the "
                                                             "slowest rank's drain sets the wall clock; at least 3)")
    ap.add_argument("--budget-s", type=float, default=555.0, help="wall-clock budget: optional blocks are shortened / skipped beyond it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kpop", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-blocks", action="store_true", help="skip the c1 / c2 / c3 blocks")
    ap.add_argument("--c5-groups", type=int, default=1024)
    ap.add_argument("--c5-steps", type=int, default=2)
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    workload = (f"C4: {args.groups} groups x {args.starts} starts = {args.groups * args.starts} Hybrid A* queries per GPU per step, "
                f"{N_GRID}x{N_GRID}x72, 96 boxes/group, launch-default params, EXACT single-pop mode")
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, workload)
        return

    import torch
    import torch.distributed as dist
    import path_planning_pkg_b200 as pp

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    deadline = T_START + args.budget_s

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def in_time(margin=0.0):
        """rank 0's view of the wall-clock budget, agreed by every rank (the optional blocks contain collectives)"""
        ok = torch.tensor([1.0 if time.time() + margin < deadline else 0.0], device="cuda")
        if world > 1:
            dist.broadcast(ok, src=0)
        return bool(ok.item() == 1.0)

    # ---- the headline batch: rank r owns groups with seeds r*G .. r*G + G - 1 (distinct work per rank) ----
    P = pp.make_params(grid_size=N_GRID, resolution=RES)
    ctx = pp.Context(P, num_groups=args.groups, device=local_rank)
    groups = build_workload(args.groups, args.starts, rank * args.groups)
    t0 = time.time()
    apply_groups(ctx, groups)
    ctx.sync()
    map_build_s = time.time() - t0
    queries, qgroups, maps = select_queries(ctx, groups)
    q = ctx.make_queries(queries, qgroups)
    nq = len(q)
    lanes = max(1, min(args.lanes, args.steps))
    # resident queries per lane: the GPU holds 16 warps per SM in all; twice its share lets a lane fill the SMs the others leave idle
    # and keeps most of the lane's memory budget for the arena its queries grow into
    hw_slots = 16 * torch.cuda.get_device_properties(local_rank).multi_processor_count
    lane_slots = args.max_slots if args.max_slots > 0 else max(256, min(hw_slots, (2 * hw_slots + lanes - 1) // lanes))
    opts = ctx.make_opts(path_cap=2048, max_slots=lane_slots)
    # end to end every step uploads its queries, and the upload orders them with a small kernel on the lane's stream: that kernel
    # needs a warp slot.  With lanes x lane_slots CTAs above what the SMs hold, a freed slot goes to the next waiting search CTA and
    # the upload waits for a whole lane to drain (the first 10-step e2e region took 145 s instead of 95 s, together with the 2D
    # fields each fresh lane computed inside it).  So the e2e batches ask for one warp slot per SM less than the GPU holds.
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def e2e_opts(batches_in_flight):
        n = args.max_slots if args.max_slots > 0 else max(32, min(lane_slots, (hw_slots - sm_count) // max(1, min(lanes, batches_in_flight))))
        return n, ctx.make_opts(path_cap=2048, max_slots=n)
    if args.e2e_steps > 0:
        e2e_steps = args.e2e_steps
    else:
        e2e_steps = max(3, min(args.steps, lanes // 2 if world == 1 else lanes // 4))
    free_b, total_b = torch.cuda.mem_get_info()
    pipe = Pipeline(ctx, lanes, q, opts, pp, fetch_lanes=max(args.warmup, e2e_steps))
    pipe.set_budget(int(free_b * 0.88 / lanes))      # pools + arena of every lane; the rest stays free for NCCL and the allocator
    pipe.upload_all()       # every lane allocates its pools and computes its groups' 2D fields (launch order) before anything is timed

    # ---- warm-up, then end to end through the C ABI: pinned host queries in, results + paths + curvature out, EVERY step.  The W
    # warm-up steps run the same way (cold kernels, cold arenas, pinned buffers touched) and are collected completely before the
    # e2e clock starts, so the e2e region holds exactly its own E batches: ramp-up, bulk and the full drain of the last one.
    # (Until this revision warm-up and e2e were one stream and the clock started while the warm-up batches were still running:
    # the region did the work of W + E batches and was credited with E.) ----
    barrier()
    e2e_slots, pipe.opts = e2e_opts(e2e_steps)
    t_w0 = time.perf_counter()
    pipe.run(args.warmup, True)
    torch.cuda.synchronize()
    t_warm = time.perf_counter() - t_w0
    # wall-clock guard (SCALE runs have a per-N limit): the warm-up just showed what W batches cost on this rank, drain included;
    # if E more of them plus the K timed steps would not fit the budget, the e2e region shrinks to 3 steps -- on every rank
    # (N = 1 fits with room to spare: 445 s with 10 e2e steps; the 0.75 is the drain's share of the warm-up, which does not grow
    # with the number of batches in flight)
    if args.e2e_steps <= 0 and e2e_steps > 3 and world > 1:
        per_step = 0.75 * t_warm / max(args.warmup, 1)
        fits = elapsed() + per_step * (e2e_steps + args.steps) + 60.0 < args.budget_s
        flag = torch.tensor([1.0 if fits else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() != 1.0:
            e2e_steps = 3
    e2e_slots, pipe.opts = e2e_opts(e2e_steps)
    barrier()
    t_mark = time.perf_counter()
    pipe.run(e2e_steps, True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t_mark) * 1e3
    r2 = pipe.results(0).copy()
    # ---- device-resident timing (value): exactly K steps, queries already in HBM, CUDA events on the context's stream ----
    pipe.opts = opts
    pipe.upload_all()
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = sum(l.kernel_launches() for l in pipe.lanes)
    value_ms, value_wall = timed(ctx, barrier, lambda: pipe.run(args.steps, False))
    clocks = sampler.finish()
    timed_launches = sum(l.kernel_launches() for l in pipe.lanes) - l0
    launch_ms = list(pipe.kernel_ms)
    retried = sum(l.batch_retried() for l in pipe.lanes)
    res, _, _ = ctx.batch_fetch()
    pops = int(res["n_pops"].sum())
    assert int(r2["n_pops"].sum()) == pops, "e2e pass expanded a different number of nodes"
    h2d, d2h = pipe.h2d, pipe.d2h
    # one batch alone on an idle GPU (latency of a step without overlap), when the budget allows
    lat_ms = float("nan")
    if world == 1 and in_time(margin=150.0):        # N > 1: the slowest rank's longest query (80 s and more) is not worth the wall clock
        lat_ms, _ = timed(ctx, barrier, lambda: pipe.run(1, False))
    pipe.close()

    ms_t = torch.tensor([value_ms, e2e_ms, lat_ms], dtype=torch.float64, device="cuda")
    cnt_t = torch.tensor([float(pops), float(nq), float(res["success"].sum()), float((res["status"] != 0).sum()),
                          float(res["n_pops_bin_oob"].sum()), float((res["n_pops_bin_oob"] > 0).sum()), float(retried)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt_t, op=dist.ReduceOp.SUM)
    max_ms, max_e2e_ms, max_lat_ms = [float(v) for v in ms_t.tolist()]
    all_pops, all_q, all_succ, all_flags, all_oob, all_oob_q, all_retried = [float(v) for v in cnt_t.tolist()]
    value = all_pops * args.steps / (max_ms * 1e-3)
    e2e_value = all_pops * e2e_steps / (max_e2e_ms * 1e-3)

    peak, peak_src = peaks()
    line = None
    if rank == 0:
        achieved = pops * args.steps * ALGO_BYTES_PER_EXPANSION / (value_ms * 1e-3) / 1e9
        line = {
            "metric": "hybrid_astar_node_expansions_per_s", "value": value, "unit": "expansions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "queries_per_s": all_q * args.steps / (max_ms * 1e-3),
            "batch_latency_ms": None if max_lat_ms != max_lat_ms else max_lat_ms,
            "expansions_per_step": int(all_pops), "queries_per_step": int(all_q),
            "success_rate": all_succ / all_q, "capacity_flags": int(all_flags), "retried_queries": int(all_retried),
            "expansions_bin_oob": int(all_oob), "queries_with_bin_oob": int(all_oob_q),
            "config": {"workload": workload, "per_rank_batch": "rank r runs the groups with seeds r*64 .. r*64+63 (distinct queries per rank)",
                       "step_submission": f"steps are independent batches, submitted round-robin over {lanes} lane contexts (own stream + scratch, "
                                          "shared maps) with one synchronisation at the end: the drain of batch k overlaps batch k+1 "
                                          "(continuous batching); batch_latency_ms = one batch alone on an idle GPU",
                       "pools": "per-query containers start at 8192 closed / 4096 open / 2048 2D-open entries and grow x2 from the lane's arena "
                                f"(library defaults); {lane_slots} resident queries per lane ({e2e_slots} in the e2e region, which leaves "
                                "the upload kernels a warp slot per SM); retried_queries = re-executions after a capacity miss",
                       "l2": "per-query scratch (open / closed sets, lazy-A* cache) of the resident queries is tens of GB, far larger than the 126 MB L2",
                       "map_build_s": map_build_s, "timed_region_wall_s": value_wall},
            "e2e": {"value": e2e_value, "unit": "expansions/s", "steps": e2e_steps, "ms_per_step": max_e2e_ms / e2e_steps,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "timed": "host clock from the submission of the first e2e batch (GPU idle, warm-up collected) to the collection of "
                             "the last one: ramp-up and the full drain of the longest query are inside the region"},
            "gpu_launches": int(timed_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": committed_traffic("pp_search_kernel"), "kernel": "pp_search_kernel", "peak_source": peak_src,
                         "launch_ms_mean": float(np.mean(launch_ms)) if launch_ms else None, "launches_in_flight": lanes,
                         "note": "latency-bound pointer chasing (libstdc++-exact red-black-tree walks on one control lane per query); "
                                 "achieved = algorithmic bytes (312 B/expansion, SURVEY 8d) of the timed region / its duration"},
            "clocks": clocks,
        }

    # ---- CPU baseline (rank 0, N = 1 only): the unmodified reference on the host cores, bounded sample of the same batch ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import orc
            n_threads = os.cpu_count() or 1
            n_s = args.cpu_sample if args.cpu_sample > 0 else int(min(max(8 * n_threads, 64), 256))
            idx = reference_sample(nq, n_s)
            Pr = orc.make_params(grid_size=N_GRID, resolution=RES)
            b = orc.ref_batch(Pr, groups, queries, qgroups, maps, idx, n_threads)
            cpu = (idx, b)
            rs = pipe.results(0)
            defined = (b["pops_oob"] == 0) & (res["n_pops_bin_oob"][idx] == 0)
            same_n = (b["pops"] == res["n_pops"][idx]) & (b["success"] == res["success"][idx])
            same_c = b["cost"].view(np.uint32) == res["cost"][idx].view(np.uint32)
            hp = np.frombuffer(pipe.hpath[0].numpy().tobytes(), np.float32).reshape(nq, 2048, 3)
            hc = np.frombuffer(pipe.hcurv[0].numpy().tobytes(), np.float32).reshape(nq, 2048)
            same_h = np.array([(not rs["success"][k]) or orc.path_hash(hp[k, :rs["n_path"][k]], hc[k, :rs["n_path"][k]]) == int(b["hash"][j])
                               for j, k in enumerate(idx)])
            ident = int((same_n & same_c & same_h & defined).sum())
            line["cpu_baseline"] = {
                "value": float(b["pops"].sum() / b["secs"]), "unit": "expansions/s", "cores": n_threads, "kind": "reference",
                "sample": f"{len(idx)} of the {nq} queries (fixed permutation, seed 1), one reference planner per thread, scrubbed per query ({b['secs']:.1f} s)",
                "queries_per_s": len(idx) / b["secs"],
                "busy_time_expansions_per_s_per_core": float(b["pops"].sum() / b["busy_s"].sum()),
                "identical_to_gpu": {"count_cost_path_hash": f"{ident}/{int(defined.sum())}",
                                     "excluded_bin72_undefined_in_reference": int((~defined).sum())}}
        except Exception as e:  # the reference .so is test infrastructure; report rather than die
            line["cpu_baseline"] = {"error": repr(e)}

    # ---- side blocks (never allowed to break the headline line) ----
    def guarded(name, fn):
        try:
            return fn()
        except Exception as e:
            return {"error": repr(e)}

    if not args.no_kpop and in_time(margin=15.0):
        k = guarded("kpop", lambda: block_kpop(args, pp, ctx, q, res, barrier, cpu, in_time()))
        if rank == 0:
            line["kpop"] = k
    ctx.close()
    if not args.no_c5 and in_time(margin=30.0):
        c5 = guarded("c5", lambda: block_c5(args, pp, torch, dist, rank, local_rank, world, barrier, in_time))
        if rank == 0:
            line["c5"] = c5
    if rank == 0 and world == 1 and not args.no_blocks:
        if time.time() < deadline:
            line["c1"] = guarded("c1", lambda: block_c1(args, pp, local_rank))
        if time.time() < deadline:
            line["c2_map_update"] = guarded("c2", lambda: block_c2(args, pp, local_rank, peak))
        if time.time() < deadline:
            line["c3_fields"] = guarded("c3", lambda: block_c3(args, pp, local_rank, peak))
    if rank == 0:
        line["wall_s"] = elapsed()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
