#!/usr/bin/env python
"""bench.py -- batched Hybrid A* queries on synthetic clutter maps (BASELINE.json configs[3] / [4]).

One "step" = one pass of the hot path over one batch: 64 (map, goal) groups x 64 start poses = 4096
independent HybridAStar::find_path queries on 512 x 512 x 72 maps with 96 box obstacles each (SURVEY.md
§8d C4), EXACT single-pop mode (expansion sequence identical to the reference).  With --gpus N every rank
runs the same 4096-query batch (weak scaling, no data-path collective; every rank rasterises its maps itself,
the NCCL map broadcast is timed separately as `map_broadcast_ms`).  --workload c5 runs BASELINE configs[4]
(65 536 queries sharded by group, K-POP(32), strong scaling).

  value : node expansions / s, whole job, inputs resident in HBM, kernel timed with CUDA events
  e2e   : the same through pp_find_path_batch with pinned HOST buffers (H2D queries, D2H results+paths)
  --impl reference : the unmodified reference (oracle/_ref) on the host cores, bounded sample per step
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


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


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


def apply_groups(ctx, groups):
    """update_goal + 4 x (boxes, decay) per group through the C ABI; returns final maps and selected queries."""
    import path_planning_pkg_b200 as pp
    thr = ctx.consts().log_threshold
    queries, qgroups, maps = [], [], []
    for gi, sc in enumerate(groups):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        for _ in range(sc["rounds"]):
            ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
            ctx.decay(group=gi)
        m = ctx.get_map(gi)
        maps.append(m)
        cand = sc["start_candidates"]
        st = ctx.set_start(ctx.make_queries(cand, [gi] * len(cand)))
        free = m[st["ci"], st["cj"]] < thr
        sel = cand[free][:sc["n_starts"]]
        queries.append(sel); qgroups += [gi] * len(sel)
    return np.concatenate(queries), np.array(qgroups, np.int32), maps


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_reference_run(groups, queries, qgroups, maps, sample_idx, n_threads):
    """The unmodified reference (oracle/_ref, stock libm) on the host cores over `sample_idx` queries."""
    import orc
    lib = C.CDLL(orc.REF_SO)
    lib.ref_bench_queries.restype = C.c_double
    P = orc.make_params(grid_size=N_GRID, resolution=RES)
    G = len(groups)
    frames = np.zeros((G, 6), np.float32)
    nb = len(groups[0]["boxes"])
    boxes = np.zeros((G, nb, 4), np.float32); conf = np.zeros((G, nb), np.float32)
    for g, sc in enumerate(groups):
        frames[g, :3] = sc["goal"]; frames[g, 3:] = sc["frame_start"]
        boxes[g] = sc["boxes"]; conf[g] = sc["conf"]
    mp = np.ascontiguousarray(np.stack(maps), np.float32)
    q4 = np.ascontiguousarray(queries[sample_idx], np.float32)
    go = np.ascontiguousarray(qgroups[sample_idx], np.int32)
    n = len(q4)
    cost = np.zeros(n, np.float32); succ = np.zeros(n, np.int32); pops = np.zeros(n, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    secs = lib.ref_bench_queries(C.byref(P), vp(frames), vp(mp), vp(boxes), vp(conf), C.c_int(nb), C.c_float(S.APF_ADDED_RADIUS),
                                 C.c_int(G), vp(q4), vp(go), C.c_int(n), C.c_int(n_threads), vp(cost), vp(succ), vp(pops))
    return secs, pops, cost, succ


def run_c5(args, rank, local_rank, world):
    """BASELINE configs[4] / SURVEY §8d C5: 65 536 queries = 1 024 (map, goal) groups x 64 starts, sharded by group
    round-robin over the ranks (STRONG scaling: the total is fixed), K-POP(32) mode.  Every rank rasterises the maps of
    its own groups (the bit-exact kernel makes replicas identical); the NCCL map broadcast is timed separately."""
    import torch
    import torch.distributed as dist
    import path_planning_pkg_b200 as pp

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total_groups = args.c5_groups
    # --c5-shard group: rank r owns groups r, r+world, ... (fields only for its own groups, but the ranks' work
    #                   differs by the difficulty of their groups)
    # --c5-shard query (default): every rank holds every map (replicated, as after the broadcast) and takes every world-th
    #                   query of every group: equal work per rank
    mine = list(range(rank, total_groups, world)) if args.c5_shard == "group" else list(range(total_groups))
    P = pp.make_params(grid_size=N_GRID, resolution=RES)
    ctx = pp.Context(P, num_groups=len(mine), device=local_rank)
    groups = [S.c4_group(g, n_starts=args.starts, grid_size=N_GRID, resolution=RES) for g in mine]
    t0 = time.time()
    queries, qgroups, _ = apply_groups(ctx, groups)
    map_build_s = time.time() - t0
    map_bcast_ms = None
    if world > 1:
        scratch = torch.empty(N_GRID * N_GRID, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            dist.broadcast(scratch, src=0)
        e1.record(); torch.cuda.synchronize()
        map_bcast_ms = e0.elapsed_time(e1) / 8
    if args.c5_shard == "query":
        from path_planning_pkg_b200.shard import shard_queries
        mine_q = shard_queries(len(queries), rank, world)
        queries, qgroups = queries[mine_q], qgroups[mine_q]
    q = ctx.make_queries(queries, qgroups)
    nq = len(q)
    pc = 1024
    opts = ctx.make_opts(max_expansions=1 << 18, path_cap=pc, max_slots=args.max_slots, mode=1, kpop=32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync()

    ctx.batch_upload(q, opts)
    for _ in range(args.warmup):
        ctx.batch_run()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    l0 = ctx.kernel_launches()
    ms = [ctx.batch_run() for _ in range(args.steps)]
    barrier()
    clocks = sampler.finish()
    launches = ctx.kernel_launches() - l0
    retried = ctx.batch_retried()
    res, _, _ = ctx.batch_fetch()
    pops = int(res["n_pops"].sum())
    # end to end: host buffers in, results + paths out
    hq = torch.from_numpy(q.view(np.uint8).copy()).pin_memory()
    hres = torch.zeros(nq * pp._cabi.RESULT_DT.itemsize, dtype=torch.uint8).pin_memory()
    hpath = torch.zeros(nq * pc * 3, dtype=torch.float32).pin_memory()
    hcurv = torch.zeros(nq * pc, dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = ctx.lib.pp_find_path_batch(ctx.h, C.c_void_p(hq.data_ptr()), C.c_int(nq), C.byref(opts), C.c_void_p(hres.data_ptr()),
                                        C.c_void_p(hpath.data_ptr()), C.c_void_p(hcurv.data_ptr()), None)
        if rc != 0:
            raise RuntimeError(ctx.lib.pp_last_error().decode())
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([float(np.sum(ms)), e2e_s], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(pops), float(nq), float(res["success"].sum()), float((res["status"] != 0).sum()), float(retried)],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        max_ms, e2e_max = float(t[0].item()), float(t[1].item())
        all_pops, all_q, all_succ, all_flags, all_retried = [float(v) for v in cnt.tolist()]
        peak, peak_src = peaks()
        achieved = all_pops / world * ALGO_BYTES_PER_EXPANSION / (max_ms / args.steps * 1e-3) / 1e9
        line = {"metric": "hybrid_astar_node_expansions_per_s", "value": all_pops * args.steps / (max_ms * 1e-3), "unit": "expansions/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "queries_per_s": all_q * args.steps / (max_ms * 1e-3), "expansions_per_step": int(all_pops), "queries_per_step": int(all_q),
                "success_rate": all_succ / all_q, "capacity_flags": int(all_flags), "retried_queries": int(all_retried),
                "config": {"workload": f"C5: {total_groups} groups x {args.starts} starts = {int(all_q)} Hybrid A* queries in total, sharded by {args.c5_shard} over "
                                       f"{world} GPU(s), {N_GRID}x{N_GRID}x72, 96 boxes/group, K-POP(32) mode (own semantics, DESIGN.md section 9)",
                           "l2": "per-query pools are tens of GB per step, far larger than the 126 MB L2",
                           "map_build_s": map_build_s, "map_broadcast_ms": map_bcast_ms},
                "e2e": {"value": all_pops * e2e_steps / e2e_max, "unit": "expansions/s", "queries_per_s": all_q * e2e_steps / e2e_max,
                        "h2d_bytes_per_step": int(q.nbytes) * world,
                        "d2h_bytes_per_step": int(hres.numel() + hpath.numel() * 4 + hcurv.numel() * 4) * world},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "kernel": "pp_kpop_kernel<4>", "peak_source": peak_src,
                             "note": "latency / barrier bound (DESIGN.md section 9); algorithmic bytes = 312 B/expansion (SURVEY 8d)"},
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--groups", type=int, default=64)
    ap.add_argument("--starts", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=64, help="queries in the cpu_baseline sample")
    ap.add_argument("--max-slots", type=int, default=0)
    # EXACT-mode pools: the library defaults (131 072 closed states per query, automatic re-run of the queries that need more in 8x
    # larger pools).  Measured structure of a step (profiles/r1_bench_search_launches.csv): the main launch on 2 368 slots = 11.4 s,
    # then the retry launch of the 116 long queries = 28.1 s, bounded by the batch's longest query.  Larger first-pass pools
    # (--max-expansions 1048576 --max-open 524288 --max-open2d 65536 --exact-slots N) avoid the re-run but leave room for fewer
    # resident queries; which side wins has to be measured (DESIGN.md sections 7 and 13).
    ap.add_argument("--max-expansions", type=int, default=1 << 17)
    ap.add_argument("--max-open", type=int, default=1 << 16)
    ap.add_argument("--max-open2d", type=int, default=1 << 14)
    ap.add_argument("--exact-slots", type=int, default=0, help="resident EXACT-mode query slots (0 = auto); --max-slots overrides")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=1, help="timed end-to-end steps (each is a full batch)")
    ap.add_argument("--no-kpop", action="store_true", help="skip the additional K-POP(32) throughput measurement")
    ap.add_argument("--no-kpop-large", action="store_true", help="skip the K-POP(32) measurement on the C5-sized batch")
    ap.add_argument("--large-starts", type=int, default=1024, help="starts per group of the C5-sized K-POP batch")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"], help="c4: EXACT-mode headline (default); c5: 65 536 queries, K-POP(32), strong scaling")
    ap.add_argument("--c5-groups", type=int, default=1024)
    ap.add_argument("--c5-shard", default="query", choices=["query", "group"])
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.workload == "c5" and args.impl == "b200":
        return run_c5(args, rank, local_rank, world)
    n_threads = os.cpu_count() or 1
    workload = (f"C4: {args.groups} groups x {args.starts} starts = {args.groups * args.starts} Hybrid A* queries per GPU, "
                f"{N_GRID}x{N_GRID}x72, 96 boxes/group, launch-default params, EXACT single-pop mode")

    if args.impl == "reference":
        if rank != 0:
            return
        import orc
        groups = build_workload(args.groups, args.starts, 0)
        # maps and start selection come from the reference itself here (no GPU on this arm)
        P = orc.make_params(grid_size=N_GRID, resolution=RES)
        maps, queries, qgroups = [], [], []
        o = orc.ref(P)
        for gi, sc in enumerate(groups):
            o.set_map(np.zeros((N_GRID, N_GRID), np.float32))
            S.build_map(o, sc)
            m = o.get_map(); maps.append(m)
            sel = S.select_starts(sc, m, o.consts().log_threshold, o.set_start)
            queries.append(sel); qgroups += [gi] * len(sel)
        queries = np.concatenate(queries); qgroups = np.array(qgroups, np.int32)
        per_step = max(n_threads * 2, 32)
        rs = np.random.RandomState(0)
        order = rs.permutation(len(queries))
        tot_pops, tot_s, tot_q, k = 0, 0.0, 0, 0
        for step in range(args.warmup + args.steps):
            idx = order[(k * per_step) % len(order):][:per_step]; k += 1
            secs, pops, _, _ = cpu_reference_run(groups, queries, qgroups, maps, idx, n_threads)
            if step >= args.warmup:
                tot_pops += int(pops.sum()); tot_s += secs; tot_q += len(idx)
        val = tot_pops / tot_s
        line = {"impl": "reference", "metric": "hybrid_astar_node_expansions_per_s", "value": val, "unit": "expansions/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "queries_per_s": tot_q / tot_s,
                "config": {"workload": workload, "sample": f"{per_step} queries per step drawn from the 4096-query batch"},
                "cpu_baseline": {"value": val, "unit": "expansions/s", "cores": n_threads, "kind": "reference",
                                 "sample": f"{per_step} queries/step x {args.steps} steps, one HybridAStar<float> per thread, scrubbed per query"},
                "e2e": {"value": val, "unit": "expansions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import path_planning_pkg_b200 as pp

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    P = pp.make_params(grid_size=N_GRID, resolution=RES)
    ctx = pp.Context(P, num_groups=args.groups, device=local_rank)
    # Weak scaling: every rank runs the SAME 4096-query batch (same seeds).  The step time of this batch is the latency of
    # its single longest query; per-rank batches of different seeds would turn the max-over-ranks
    # time into an extreme-value statistic of that one query instead of a measurement of scaling.
    groups = build_workload(args.groups, args.starts, 0)
    t0 = time.time()
    queries, qgroups, maps = apply_groups(ctx, groups)
    map_build_s = time.time() - t0

    # map replication (north_star: "map replicated by an NCCL broadcast over NVLink after each update"):
    # timed separately; every rank re-broadcasts its first group's map from rank 0's buffer and back-checks size only
    map_bcast_ms = None
    if world > 1:
        class _Alias:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        scratch = torch.empty(N_GRID * N_GRID, dtype=torch.float32, device="cuda")
        src = torch.as_tensor(_Alias(ctx.map_device_ptr(0), N_GRID * N_GRID), device="cuda")
        scratch.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            dist.broadcast(scratch, src=0)
        e1.record(); torch.cuda.synchronize()
        map_bcast_ms = e0.elapsed_time(e1) / 8

    q = ctx.make_queries(queries, qgroups)
    nq = len(q)
    exact_slots = args.max_slots if args.max_slots > 0 else args.exact_slots
    opts = ctx.make_opts(max_expansions=args.max_expansions, max_open=args.max_open, max_open2d=args.max_open2d, path_cap=2048,
                         max_slots=exact_slots)
    pool_note = "library defaults" if args.max_expansions == 1 << 17 else "caller-sized pools"
    try:
        ctx.batch_upload(q, opts)
    except pp.PPError as e:      # e.g. not enough free memory on this device: the library defaults (small pools, automatic 8x retries)
        pool_note = f"library defaults after the sized pools could not be set up ({e})"
        args.max_expansions, args.max_open, args.max_open2d = 1 << 17, 1 << 16, 1 << 14
        opts = ctx.make_opts(max_expansions=args.max_expansions, max_open=args.max_open, max_open2d=args.max_open2d, path_cap=2048,
                             max_slots=args.max_slots)
    launches0 = ctx.kernel_launches()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync()

    # ---- device-resident timing (value) ----
    ctx.batch_upload(q, opts)
    for _ in range(args.warmup):
        ctx.batch_run()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    kernel_ms = []
    launches1 = ctx.kernel_launches()
    for _ in range(args.steps):
        kernel_ms.append(ctx.batch_run())
    barrier()
    clocks = sampler.finish()
    timed_launches = ctx.kernel_launches() - launches1
    retried_exact = ctx.batch_retried()
    res, _, _ = ctx.batch_fetch()
    pops = int(res["n_pops"].sum())
    total_ms = float(np.sum(kernel_ms))
    ms_t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    pops_t = torch.tensor([pops * args.steps, nq * args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pops_t, op=dist.ReduceOp.SUM)
    max_ms = float(ms_t.item()); all_pops, all_q = [float(v) for v in pops_t.tolist()]
    value = all_pops / (max_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers ----
    pc = 2048
    hq = torch.from_numpy(q.view(np.uint8).copy()).pin_memory()
    hres = torch.zeros(nq * pp._cabi.RESULT_DT.itemsize, dtype=torch.uint8).pin_memory()
    hpath = torch.zeros(nq * pc * 3, dtype=torch.float32).pin_memory()
    hcurv = torch.zeros(nq * pc, dtype=torch.float32).pin_memory()
    lib = ctx.lib

    def e2e_step():
        rc = lib.pp_find_path_batch(ctx.h, C.c_void_p(hq.data_ptr()), C.c_int(nq), C.byref(opts), C.c_void_p(hres.data_ptr()),
                                    C.c_void_p(hpath.data_ptr()), C.c_void_p(hcurv.data_ptr()), None)
        if rc != 0:
            raise RuntimeError(lib.pp_last_error().decode())

    # (the kernel and the pools are warm from the device-resident steps above: no extra warm-up pass)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = (all_pops / args.steps * e2e_steps) / float(e2e_t.item())
    r2 = np.frombuffer(hres.numpy().tobytes(), pp._cabi.RESULT_DT)
    assert int(r2["n_pops"].sum()) == pops, "e2e pass expanded a different number of nodes"

    # ---- K-POP(32) throughput mode on the same batch (new semantics: results are the K-POP restatement's, not the
    # reference's; reported beside the headline, never instead of it) ----
    kpop_info = None
    if not args.no_kpop:
        kopts = ctx.make_opts(max_expansions=1 << 18, path_cap=2048, max_slots=args.max_slots, mode=1, kpop=32)
        ctx.batch_upload(q, kopts)
        ksteps = max(args.steps, 5)
        for _ in range(args.warmup):
            ctx.batch_run()
        barrier()
        kl0 = ctx.kernel_launches()
        kms = [ctx.batch_run() for _ in range(ksteps)]
        barrier()
        k_launches = ctx.kernel_launches() - kl0
        kres, _, _ = ctx.batch_fetch()
        # end to end through the C ABI, host buffers, same as the exact-mode e2e above
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            rc = lib.pp_find_path_batch(ctx.h, C.c_void_p(hq.data_ptr()), C.c_int(nq), C.byref(kopts), C.c_void_p(hres.data_ptr()),
                                        C.c_void_p(hpath.data_ptr()), C.c_void_p(hcurv.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(lib.pp_last_error().decode())
        barrier()
        ke2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        kt = torch.tensor([float(np.sum(kms))], dtype=torch.float64, device="cuda")
        kp = torch.tensor([float(kres["n_pops"].sum()) * ksteps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ke2e, op=dist.ReduceOp.MAX)
            dist.all_reduce(kp, op=dist.ReduceOp.SUM)
        both = (res["success"] == 1) & (kres["success"] == 1)
        ratio = kres["cost"][both] / res["cost"][both] - 1.0
        kq = all_q / args.steps * ksteps
        kpop_info = {"k": 32, "expansions_per_s": float(kp.item()) / (float(kt.item()) * 1e-3),
                     "queries_per_s": kq / (float(kt.item()) * 1e-3), "ms_per_step": float(kt.item()) / ksteps, "steps": ksteps,
                     "e2e": {"expansions_per_s": float(kp.item()) / float(ke2e.item()), "queries_per_s": kq / float(ke2e.item())},
                     "gpu_launches": int(k_launches),
                     "expansions_per_step": int(kres["n_pops"].sum()), "success_rate": float(kres["success"].mean()),
                     "cost_vs_exact_mode": {"median": float(np.median(ratio)), "min": float(ratio.min()), "max": float(ratio.max())},
                     "note": "k pops per iteration, exact 2D field heuristic, no equal-f drops; bit-identical to its CPU restatement "
                             "(oracle/port/kpop.inc), NOT to the reference (SURVEY F4/F5)"}

        klat = []
        for k in range(min(16, nq)):
            t1 = time.perf_counter()
            ctx.find_path_batch(q[k:k + 1], ctx.make_opts(max_expansions=1 << 18, path_cap=pc, max_slots=1, mode=1, kpop=32))
            klat.append((time.perf_counter() - t1) * 1e3)
        kpop_info["p50_single_query_ms"] = float(np.median(klat))
        # the same mode on a C5-sized batch (BASELINE configs[4]: 65536 queries, k-pop = 32): 1024 starts on each of the
        # groups already on the device; device-resident timing only
        if not args.no_kpop_large:
            big_q, big_g = [], []
            thr = ctx.consts().log_threshold
            for gi in range(args.groups):
                cand = S.c4_group(gi, n_starts=args.large_starts, grid_size=N_GRID, resolution=RES)["start_candidates"]
                st = ctx.set_start(ctx.make_queries(cand, [gi] * len(cand)))
                free = maps[gi][st["ci"], st["cj"]] < thr
                sel = cand[free][:args.large_starts]
                big_q.append(sel); big_g += [gi] * len(sel)
            bq = ctx.make_queries(np.concatenate(big_q), np.array(big_g, np.int32))
            bopts = ctx.make_opts(max_expansions=1 << 18, path_cap=1024, max_slots=args.max_slots, mode=1, kpop=32)
            ctx.batch_upload(bq, bopts)
            for _ in range(args.warmup):
                ctx.batch_run()
            barrier()
            bms = [ctx.batch_run() for _ in range(ksteps)]
            barrier()
            bres, _, _ = ctx.batch_fetch(want_paths=False)
            bt = torch.tensor([float(np.sum(bms))], dtype=torch.float64, device="cuda")
            bp = torch.tensor([float(bres["n_pops"].sum()) * ksteps, float(len(bq)) * ksteps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(bt, op=dist.ReduceOp.MAX)
                dist.all_reduce(bp, op=dist.ReduceOp.SUM)
            kpop_info["c5_batch"] = {"queries_per_gpu": int(len(bq)), "expansions_per_s": float(bp[0].item()) / (float(bt.item()) * 1e-3),
                                     "queries_per_s": float(bp[1].item()) / (float(bt.item()) * 1e-3),
                                     "ms_per_step": float(bt.item()) / ksteps, "success_rate": float(bres["success"].mean()),
                                     "capacity_flags": int((bres["status"] != 0).sum())}

    # single-query latency (p50) on the first 16 queries, one at a time through the same ABI
    lat = []
    for k in range(min(16, nq)):
        t1 = time.perf_counter()
        ctx.find_path_batch(q[k:k + 1], ctx.make_opts(max_expansions=args.max_expansions, max_open=args.max_open,
                                                      max_open2d=args.max_open2d, path_cap=pc, max_slots=1))
        lat.append((time.perf_counter() - t1) * 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    avg_launch_s = (total_ms / args.steps) * 1e-3
    achieved = (pops * ALGO_BYTES_PER_EXPANSION) / avg_launch_s / 1e9
    line = {
        "metric": "hybrid_astar_node_expansions_per_s", "value": value, "unit": "expansions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "queries_per_s": all_q / (max_ms * 1e-3),
        "p50_single_query_ms": float(np.median(lat)),
        "expansions_per_step": pops, "queries_per_step": nq,
        "success_rate": float(res["success"].mean()), "capacity_flags": int((res["status"] != 0).sum()),
        "expansions_bin_oob": int(res["n_pops_bin_oob"].sum()),
        "config": {"workload": workload, "slots": int(opts.max_slots) or "auto",
                   "pools": {"max_expansions": int(opts.max_expansions), "max_open": int(opts.max_open), "max_open2d": int(opts.max_open2d),
                             "note": pool_note, "retried_queries": int(retried_exact)},
                   "per_rank_batch": "identical on every rank (same seeds): the step is bound by the batch's single longest query",
                   "l2": "per-query scratch (open/closed sets, lazy-A* cache) is tens of GB per step, far larger than the 126 MB L2",
                   "map_build_s": map_build_s, "map_broadcast_ms": map_bcast_ms},
        "e2e": {"value": e2e_value, "unit": "expansions/s", "h2d_bytes_per_step": int(q.nbytes),
                "d2h_bytes_per_step": int(hres.numel() + hpath.numel() * 4 + hcurv.numel() * 4)},
        "gpu_launches": int(timed_launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "pp_search_kernel", "peak_source": peak_src,
                     "note": "latency-bound pointer chasing (libstdc++-exact rb-tree walks); algorithmic bytes = 312 B/expansion (SURVEY 8d)"},
        "clocks": clocks,
        "kpop": kpop_info,
    }
    # north_star (c): the generic vehicle-footprint collision kernel on this batch's first map (4.0 x 2.0 m rectangle, 2^20 poses);
    # a side measurement, never allowed to break the headline line
    try:
        rs = np.random.RandomState(3)
        L = N_GRID * RES
        fp = np.concatenate([rs.uniform(0.0, L, (1 << 20, 2)), rs.uniform(-np.pi, np.pi, (1 << 20, 1))], 1).astype(np.float32)
        ctx.set_footprint(4.0, 2.0, 1.0)
        ctx.footprint(fp)
        fms = min(ctx.footprint(fp, want_ms=True)[3] for _ in range(5))
        fcells = float(np.mean([len(ctx.footprint_table(b)) for b in range(72)]))
        line["footprint_kernel"] = {"poses": len(fp), "vehicle_m": [4.0, 2.0, 1.0], "mean_cells_per_pose": fcells, "kernel_ms": fms,
                                    "poses_per_s": len(fp) / (fms * 1e-3),
                                    "algorithmic_GBps": len(fp) * (fcells * 4 + 24) / (fms * 1e-3) / 1e9,
                                    "note": "map is L1/L2-resident (1 MiB): issue/L1-bound, see DESIGN.md section 12"}
    except Exception as e:
        line["footprint_kernel"] = {"error": str(e)}
    if not args.no_cpu_baseline:
        try:
            rs = np.random.RandomState(1)
            idx = rs.permutation(nq)[:args.cpu_sample]
            secs, cpops, ccost, _ = cpu_reference_run(groups, queries, qgroups, maps, idx, n_threads)
            same = int((cpops == res["n_pops"][idx]).sum())
            line["cpu_baseline"] = {"value": float(cpops.sum() / secs), "unit": "expansions/s", "cores": n_threads, "kind": "reference",
                                    "sample": f"{len(idx)} of the {nq} queries, one reference planner per thread, scrubbed per query ({secs:.1f} s)",
                                    "queries_per_s": len(idx) / secs,
                                    "same_expansion_count_as_gpu": f"{same}/{len(idx)} (stock glibc libm vs pinned libm, see DESIGN.md)"}
            # also the single-threaded reference as it ships (one planner, one core), on the 8 shortest queries
            # of the same sample so that the side measurement stays bounded
            try:
                short = idx[np.argsort(cpops)[:8]]
                s1, p1, _, _ = cpu_reference_run(groups, queries, qgroups, maps, short, 1)
                line["cpu_baseline"]["one_core"] = {"value": float(p1.sum() / s1), "unit": "expansions/s", "queries_per_s": len(short) / s1,
                                                    "sample": f"the {len(short)} shortest queries of the sample above ({s1:.1f} s)"}
            except Exception as e:
                line["cpu_baseline"]["one_core"] = {"error": str(e)}
        except Exception as e:  # the reference .so is test infrastructure; report rather than die
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
