"""GPU parity tests (run with `pytest -m gpu` on a B200): every call goes through the C ABI of
libpp_b200.so and is compared with the compiled, unmodified reference (oracle/_ref).

The oracle is `orc.ref`: the reference's own objects linked against the stock glibc of this image.  The device
restates glibc's binary32 sinf / cosf / atan2f / acosf (csrc/core/pp_gmath.h, pinned exhaustively by
tests/test_cpu_gmath.py), so EVERYTHING here is bit-exact against the stock build: map, indices, collision booleans,
roll-out, APF and Dubins values, the expansion sequence, cost, path and curvature.  (`crm` below is an alias of the
same stock oracle, kept so that the test bodies read as before; the pinned-libm flavour of round 1 is a build variant,
tests/test_gpu_pinned_variant.py.)
"""
import numpy as np
import pytest

import orc
import scenarios as S

pytestmark = pytest.mark.gpu


def _ctx(P, groups=1):
    import path_planning_pkg_b200 as pp
    return pp.Context(pp._cabi.params_from(P), num_groups=groups, device=0)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _states_equal(a, b):
    for f in a.dtype.names:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            if not np.array_equal(_bits(x), _bits(y)):
                return False, f
        elif not np.array_equal(x, y):
            return False, f
    return True, None


@pytest.fixture(scope="module")
def golden():
    P = orc.ref_test_params()
    ctx, ref = _ctx(P), orc.ref(P)
    crm = ref
    for o in (ctx, ref):
        orc.setup_ref_test_scenario(o)
    return P, ctx, ref, crm


def test_constants_and_tables(golden):
    P, ctx, ref, crm = golden
    c, r = ctx.consts(), ref.consts()
    for f in ("log_threshold", "log_min", "log_max", "log_free", "precision", "r_min", "ang_step"):
        assert np.float32(getattr(c, f)) == np.float32(getattr(r, f)), f
    fr = ctx.frame()
    assert np.float32(fr.grid_heading) == np.float32(r.grid_heading)
    assert list(fr.goal_grid) == list(r.goal_grid) and fr.goal_bin == r.goal_bin
    assert (fr.goal_ci, fr.goal_cj) == (r.goal_ci, r.goal_cj)
    for a, b in zip(ctx.tables(), ref.tables()):
        assert np.array_equal(_bits(a), _bits(b))


def test_map_golden_scenario_bitexact(golden):
    """decay + lane lines + boxes x5 (utils/hybrid_astar/test_hybrid_astar.cpp:77-84): all N*N floats bitwise."""
    P, ctx, ref, crm = golden
    m = ctx.get_map()
    assert np.array_equal(_bits(m), _bits(ref.get_map()))
    assert (m >= ref.consts().log_threshold).sum() > 100


def test_rollout_bitexact(golden):
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(3)
    n = 4096
    st = np.zeros(n, orc.STATE_DT)
    st["x"] = rs.uniform(0, 30, n); st["y"] = rs.uniform(0, 30, n); st["heading"] = rs.uniform(-3.05, 3.05, n)
    st["g"] = rs.uniform(0, 50, n); st["vmin_sqr"] = rs.uniform(0, 9, n)
    st["curvature_index"] = rs.randint(0, P.num_steering, n)
    prec = ref.consts().precision
    st["angle_bin"] = [int((np.float32(np.round(np.float32(h) / np.float32(prec)) * np.float32(prec)) + np.pi) / float(prec)) for h in st["heading"]]
    st["f"] = st["g"]; st["ci"] = -1; st["cj"] = -1
    assert (st["angle_bin"] < P.num_angle_bins).all()
    a, ac, af = ctx.rollout(st)
    b, bc, bf = ref.rollout(st)
    assert np.array_equal(ac, bc) and np.array_equal(af, bf)
    ok, f = _states_equal(a, b)
    assert ok, f


def test_expand_collision_apf(golden):
    """Grid3D::get_neighbors: survivors (collision booleans), cells and roll-out bit-exact vs stock reference;
    APF-augmented g bit-exact vs the pinned-libm reference and within 1e-5 of the stock one."""
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(4)
    n = 8192
    st = np.zeros(n, orc.STATE_DT)
    st["x"] = rs.uniform(0, 30, n); st["y"] = rs.uniform(0, 30, n); st["heading"] = rs.uniform(-3.05, 3.05, n)
    st["g"] = rs.uniform(0, 50, n); st["vmin_sqr"] = rs.uniform(0, 9, n)
    st["curvature_index"] = rs.randint(0, P.num_steering, n)
    prec = np.float32(ref.consts().precision)
    st["angle_bin"] = ((np.round(st["heading"] / prec).astype(np.float32) * prec).astype(np.float64) + np.pi) / float(prec)
    st["f"] = st["g"]
    a, ac, af = ctx.expand(st)
    b, bc, bf = ref.expand(st)
    c, cc, cf = crm.expand(st)
    assert np.array_equal(ac, bc) and np.array_equal(af, bf)          # same survivors = same collision booleans
    for f in ("x", "y", "heading", "vmin_sqr"):
        assert np.array_equal(_bits(a[f]), _bits(b[f])), f
    for f in ("curvature_index", "angle_bin", "ci", "cj"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(_bits(a["g"]), _bits(c["g"])), "APF cost differs from pinned-libm reference"
    assert (a["g"] != st["g"][:, None] + 0).any()
    rel = np.abs(a["g"] - b["g"]) / np.maximum(np.abs(b["g"]), 1e-6)
    assert rel.max() <= 1e-5, rel.max()


def test_collision_lookup_bitexact(golden):
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(5)
    n = 20000
    xy = np.stack([rs.uniform(-2, 32, n), rs.uniform(-2, 32, n)], 1).astype(np.float32)
    free, cells = ctx.collision(xy)
    m = ref.get_map(); thr = ref.consts().log_threshold; N = P.grid_size
    res = np.float32(P.resolution)
    ci = (xy[:, 0] / res).astype(np.int32); cj = (xy[:, 1] / res).astype(np.int32)   # trunc toward zero, float32 divide
    inb = (ci > -1) & (ci < N) & (cj > -1) & (cj < N)
    exp = np.zeros(n, bool)
    exp[inb] = m[ci[inb], cj[inb]] < thr
    assert np.array_equal(cells[:, 0], ci) and np.array_equal(cells[:, 1], cj)
    assert np.array_equal(free, exp)
    assert 0.05 < exp.mean() < 0.99
    # rounded-index variant used by the Dubins shot (Grid3D::check_path)
    for k in range(0, 400, 7):
        pts = np.concatenate([xy[k:k + 5], np.zeros((5, 1), np.float32)], 1)
        assert ctx.check_path(pts) == ref.check_path(pts)


def test_apf_values(golden):
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(6)
    n = 20000
    xyh = np.stack([rs.uniform(2, 28, n), rs.uniform(2, 28, n), rs.uniform(-3.14, 3.14, n)], 1).astype(np.float32)
    a = ctx.apf(xyh); b = crm.apf(xyh); r = ref.apf(xyh)
    assert (b > 0).sum() > 1000
    assert np.array_equal(_bits(a), _bits(b))
    assert np.array_equal(_bits(a), _bits(r))


def test_dubins_length(golden):
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(7)
    n = 50000
    goal = np.array(list(ref.consts().goal_grid), np.float32)
    starts = np.stack([rs.uniform(0, 30, n), rs.uniform(0, 30, n), rs.uniform(-3.14, 3.14, n)], 1).astype(np.float32)
    a, at, ap = ctx.dubins_length(starts, goal)
    b, bt, bp = crm.dubins_length(starts, goal)
    r, rt, rp = ref.dubins_length(starts, goal)
    assert np.array_equal(_bits(a), _bits(b)) and np.array_equal(at, bt) and np.array_equal(_bits(ap), _bits(bp))
    # north_star asks for 1e-5 relative; the glibc restatement gives the stock build's bits (NaN candidates included)
    assert np.array_equal(_bits(a), _bits(r)) and np.array_equal(at, rt) and np.array_equal(_bits(ap), _bits(rp))


def test_dubins_length_fp32(golden):
    """FP32 SIMT flavour (K-POP heuristic): bit-identical to its CPU restatement (oracle/port/fmath.inc, libm mode 2) and
    within 1e-5 relative of the reference's lengths (north_star), branch flips of the +-2pi corrections counted."""
    P, ctx, ref, crm = golden
    port = orc.port(P)
    orc.setup_ref_test_scenario(port)
    rs = np.random.RandomState(11)
    n = 200000
    goal = np.array(list(ref.consts().goal_grid), np.float32)
    starts = np.stack([rs.uniform(0, 30, n), rs.uniform(0, 30, n), rs.uniform(-3.14, 3.14, n)], 1).astype(np.float32)
    a = ctx.dubins_length_fp32(starts, goal)
    port.lib.port_set_libm(2)
    try:
        b, _, _ = port.dubins_length(starts, goal)
    finally:
        port.lib.port_set_libm(0)
    assert np.array_equal(_bits(a), _bits(b))
    r, _, _ = ref.dubins_length(starts, goal)
    rel = np.abs(a - r) / np.maximum(np.abs(r), 1e-6)
    jumps = int((rel > 1e-5).sum())
    print(f"fp32 dubins vs reference: max rel (non-jump) {rel[rel <= 1e-5].max():.3g}, branch flips {jumps}/{n}")
    assert jumps <= n // 2000


def test_dubins_path_golden(golden):
    """utils/dubins_paths.py:6 scenario: (0,0,0) -> (20,-20,pi/2)."""
    P, ctx, ref, crm = golden
    s = np.array([0, 0, 0], np.float32); g = np.array([20, -20, np.pi / 2], np.float32)
    a = ctx.dubins_path(s, g); b = crm.dubins_path(s, g)
    assert len(a[0]) == len(b[0]) and a[3] == b[3]
    assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
    assert np.float32(a[2]) == np.float32(b[2])


def test_lazy_astar_sequence(golden):
    """AStar::find_path(i,j) is history dependent: same query order -> identical values."""
    P, ctx, ref, crm = golden
    rs = np.random.RandomState(8)
    m = ref.get_map(); thr = ref.consts().log_threshold
    free = np.argwhere(m < thr)
    ij = free[rs.choice(len(free), 600, replace=False)].astype(np.int32)
    ref.scrub()
    b = ref.astar_lazy(ij)
    a = ctx.astar_lazy(ij)
    assert np.array_equal(_bits(a), _bits(b))
    assert (b < 1e30).sum() > 300


def test_search_golden_sequence_and_path(golden):
    """utils/hybrid_astar/plot.py:47-51: cost 33.0305, 43 points; full expansion sequence vs the reference."""
    P, ctx, ref, crm = golden
    ref.scrub()
    a = ctx.find_path(2.0, orc.REF_TEST_START)
    b = crm.find_path(2.0, orc.REF_TEST_START)
    assert a["success"] and a["status"] == 0
    assert abs(float(a["cost"]) - 33.0305) < 1e-4 and len(a["path"]) == 43
    assert a["n_pops"] == b["n_pops"] == 882
    ok, f = _states_equal(a["pops"], b["pops"])
    assert ok, f
    assert np.array_equal(_bits(a["path"]), _bits(b["path"])) and np.array_equal(_bits(a["curvature"]), _bits(b["curvature"]))
    assert a["cost"] == b["cost"]
    golden_path = np.load(orc.ROOT + "/tests/golden/hybrid_astar_path.npy")
    assert np.allclose(a["path"][::-1], golden_path, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("seed", list(range(8)))
def test_search_c1_exact(seed):
    sc = S.c1_scenario(seed)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, ref = _ctx(P), orc.ref(P)
    crm = ref
    for o in (ctx, ref):
        S.build_map(o, sc)
    assert np.array_equal(_bits(ctx.get_map()), _bits(ref.get_map()))
    q = sc["queries"][0]
    a = ctx.find_path(float(q[3]), q[:3]); b = crm.find_path(float(q[3]), q[:3])
    assert a["status"] == 0 and a["success"] == b["success"] and a["n_pops"] == b["n_pops"]
    ok, f = _states_equal(a["pops"], b["pops"])
    assert ok, f
    assert a["cost"] == b["cost"] and np.array_equal(_bits(a["path"]), _bits(b["path"]))


def test_search_c1_100_seeds_identical_to_stock_reference():
    """seeds 0-99 (N=200, 0.2 m, 5 boxes, 4 rounds).  Every query whose reference run is defined (no pop in
    heading bin 72) must be identical to the STOCK reference build: expansion count, cost bits, path and
    curvature bits.  Single-query latencies (reference on one host core vs the GPU through the C ABI) are reported."""
    import json
    import os
    import time
    seeds = list(range(100))
    scs = [S.c1_scenario(s) for s in seeds]
    P = orc.make_params(grid_size=scs[0]["grid_size"], resolution=scs[0]["resolution"])
    ctx = _ctx(P, groups=len(seeds))
    ref = orc.ref(P)
    queries = []
    for gi, sc in enumerate(scs):
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        for _ in range(sc["rounds"]):
            ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
            ctx.decay(group=gi)
        queries.append(sc["queries"][0])
    q = ctx.make_queries(np.array(queries), list(range(len(seeds))))
    opts = ctx.make_opts(path_cap=2048)
    res, paths, curv, _ = ctx.find_path_batch(q, opts)
    lat_gpu, lat_cpu, undefined, pops = [], [], 0, []
    for gi, sc in enumerate(scs):
        qq = sc["queries"][0]
        ref.set_map(np.zeros((sc["grid_size"], sc["grid_size"]), np.float32))
        S.build_map(ref, sc)
        assert np.array_equal(_bits(ctx.get_map(gi)), _bits(ref.get_map())), gi
        ref.scrub()                       # reset() alone leaves stale heuristic state
        t0 = time.perf_counter(); b = ref.find_path(float(qq[3]), qq[:3]); lat_cpu.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter(); ctx.find_path_batch(q[gi:gi + 1], ctx.make_opts(path_cap=2048, max_slots=1)); lat_gpu.append((time.perf_counter() - t0) * 1e3)
        r = res[gi]
        assert r["status"] == 0
        if b["n_pops_bin_oob"] > 0 or r["n_pops_bin_oob"] > 0:
            undefined += 1
            continue
        assert bool(r["success"]) == b["success"] and r["n_pops"] == b["n_pops"], (gi, r["n_pops"], b["n_pops"])
        assert np.float32(r["cost"]) == b["cost"]
        assert np.array_equal(_bits(paths[gi, :r["n_path"]]), _bits(b["path"]))
        assert np.array_equal(_bits(curv[gi, :r["n_path"]]), _bits(b["curvature"]))
        pops.append(int(r["n_pops"]))
    rep = {"config": "C1: N=200, res 0.2, 5 boxes, 4 rounds, seeds 0-99, EXACT mode",
           "identical_to_stock_glibc_reference": len(pops), "undefined_in_reference_bin72": undefined,
           "pops_p50": float(np.median(pops)), "pops_p95": float(np.percentile(pops, 95)),
           "gpu_single_query_ms": {"p50": float(np.median(lat_gpu)), "p95": float(np.percentile(lat_gpu, 95))},
           "cpu_reference_single_query_ms_1core": {"p50": float(np.median(lat_cpu)), "p95": float(np.percentile(lat_cpu, 95))}}
    print("C1 report:", json.dumps(rep))
    out = os.path.join(orc.ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(rep, open(os.path.join(out, "r2_c1_report.json"), "w"))
    assert len(pops) + undefined == 100 and undefined <= 5


def _c4_setup(seeds, n_starts):
    """C4 groups on the device and in the reference: returns ctx, params, scenario dicts, queries, their groups, the maps."""
    groups = [S.c4_group(s, n_starts=n_starts) for s in seeds]
    P = orc.make_params(grid_size=512, resolution=0.2)
    ctx = _ctx(P, groups=len(groups))
    ref = orc.ref(P)
    queries, qgroups, maps = [], [], []
    for gi, sc in enumerate(groups):
        ref.set_map(np.zeros((512, 512), np.float32))
        S.build_map(ref, sc)
        ctx.update_goal(sc["goal"], sc["frame_start"], group=gi)
        for _ in range(sc["rounds"]):
            ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, group=gi)
            ctx.decay(group=gi)
        m = ref.get_map()
        assert np.array_equal(_bits(ctx.get_map(gi)), _bits(m))
        qs = S.select_starts(sc, m, ref.consts().log_threshold, ref.set_start)
        queries += list(qs); qgroups += [gi] * len(qs); maps.append(m)
    return ctx, P, groups, np.array(queries, np.float32), np.array(qgroups, np.int32), maps


def test_search_c4_batch_exact_sequences():
    """C4 shape (512^2 x 72, 96 boxes): a batch over 2 groups x 6 starts; the full expansion SEQUENCE (every popped state with
    its g and f), cost, path and curvature of every query against the stock reference, scrubbed per query."""
    ctx, P, groups, queries, qgroups, maps = _c4_setup((0, 1), 6)
    q = ctx.make_queries(queries, qgroups)
    opts = ctx.make_opts(trace_cap=1 << 17, path_cap=2048)
    res, paths, curv, trace = ctx.find_path_batch(q, opts)
    ref = orc.ref(P)
    n_match = n_undefined = 0
    cur = -1
    for k in range(len(q)):
        if qgroups[k] != cur:
            cur = int(qgroups[k])
            ref.set_map(np.zeros((512, 512), np.float32))
            S.build_map(ref, groups[cur])
        ref.scrub()
        b = ref.find_path(float(queries[k][3]), queries[k][:3])
        r = res[k]
        assert r["status"] == 0
        if b["n_pops_bin_oob"] > 0 or r["n_pops_bin_oob"] > 0:
            # the reference indexes _offset_xy[.][72] one past the end (heap garbage) when a popped node
            # has heading bin == num_angle_bins; its results are undefined from that pop on -> counted, not compared
            n_undefined += 1
            continue
        assert bool(r["success"]) == b["success"]
        assert r["n_pops"] == b["n_pops"], (k, r["n_pops"], b["n_pops"])
        ok, f = _states_equal(trace[k, :r["n_pops"]], b["pops"])
        assert ok, (k, f)
        assert np.float32(r["cost"]) == b["cost"]
        assert np.array_equal(_bits(paths[k, :r["n_path"]]), _bits(b["path"]))
        assert np.array_equal(_bits(curv[k, :r["n_path"]]), _bits(b["curvature"]))
        n_match += 1
    print(f"C4 batch: {n_match}/{len(q)} queries identical to the stock reference, {n_undefined} undefined in the reference "
          f"(bin-72 UB), {int(res['n_pops'].sum())} expansions")
    assert n_match + n_undefined == len(q) and n_undefined <= 2


def test_search_c4_first_256_queries_identical():
    """full-set parity on the first 256 C4 queries (groups 0-3 x 64 starts).  Expansion count, cost bits,
    path length and the hash of the path + curvature bits of every query against the stock reference (all host threads, one
    planner per thread, scrubbed per query).  Queries in which the reference pops a state in heading bin 72 (undefined
    behaviour) are excluded: their number is printed and bounded."""
    ctx, P, groups, queries, qgroups, maps = _c4_setup((0, 1, 2, 3), 64)
    assert len(queries) == 256
    q = ctx.make_queries(queries, qgroups)
    res, paths, curv, _ = ctx.find_path_batch(q, ctx.make_opts(path_cap=2048))
    assert (res["status"] == 0).all()
    b = orc.ref_batch(P, groups, queries, qgroups, maps)
    excluded = (b["pops_oob"] > 0) | (res["n_pops_bin_oob"] > 0)
    bad = []
    for k in range(256):
        if excluded[k]:
            continue
        r = res[k]
        h = orc.path_hash(paths[k, :r["n_path"]], curv[k, :r["n_path"]]) if r["success"] else 0
        same = (int(r["success"]) == int(b["success"][k]) and int(r["n_pops"]) == int(b["pops"][k])
                and np.float32(r["cost"]).view(np.uint32) == b["cost"][k].view(np.uint32)
                and int(r["n_path"]) == int(b["n_path"][k]) and h == int(b["hash"][k]))
        if not same:
            bad.append((k, int(r["n_pops"]), int(b["pops"][k]), float(r["cost"]), float(b["cost"][k])))
    print(f"C4 first 256: {256 - int(excluded.sum()) - len(bad)} identical (count, cost bits, path hash), {int(excluded.sum())} excluded "
          f"(bin-72 UB in the reference), {len(bad)} different; {int(res['n_pops'].sum())} expansions; reference "
          f"{b['secs']:.1f} s on {__import__('os').cpu_count()} threads")
    assert not bad, bad[:5]
    assert excluded.sum() <= 16
