"""GPU edge cases through the C ABI (run with `pytest -m gpu`): empty and degenerate inputs, starts the reference maps
silently to cell (0, 0), capacity exhaustion and the automatic retry, output truncation flags, determinism under the
dynamic query scheduling, and invalidation of the cached 2D field after a map / goal change.  Oracles: the compiled
reference with pinned libm (EXACT mode) and the K-POP restatement (oracle/port/kpop.inc)."""
import numpy as np
import pytest

import orc
import scenarios as S
from test_gpu_kpop import port_kpop, _same
from test_gpu_parity import _bits, _ctx, _states_equal

pytestmark = pytest.mark.gpu

PATH_OVERFLOW = 8   # PP_STATUS_PATH_OVERFLOW (include/pp_b200.h)


def _status_bits():
    import re
    txt = open(orc.ROOT + "/include/pp_b200.h").read()
    return {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(PP_STATUS_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)", txt)}


def test_empty_inputs_and_free_map():
    """n = 0 boxes / lines are no-ops (Grid2D.cpp:99-139, :142-194 loop zero times); a query on the empty map is the
    reference's, expansion by expansion."""
    P = orc.ref_test_params()
    ctx, crm = _ctx(P), orc.ref(P)
    for o in (ctx, crm):
        orc.setup_ref_test_scenario(o)
    empty4 = np.zeros((0, 4), np.float32); empty1 = np.zeros(0, np.float32)
    before = ctx.get_map().copy()
    ctx.update_boxes(empty4, empty1, 1.5)
    ctx.update_lines(empty4, empty1, 0.3)
    assert np.array_equal(_bits(ctx.get_map()), _bits(before))
    # wipe the map on both sides and search on free space
    n = before.shape[0]
    for o in (ctx, crm):
        o.set_map(np.zeros((n, n), np.float32))
    crm.update_boxes(empty4, empty1, 1.5)           # clears the reference's APF list as well
    ctx.update_boxes(empty4, empty1, 1.5)
    crm.scrub()
    a = ctx.find_path(2.0, orc.REF_TEST_START)
    b = crm.find_path(2.0, orc.REF_TEST_START)
    assert a["status"] == 0 and a["success"] == b["success"] and a["n_pops"] == b["n_pops"]
    ok, f = _states_equal(a["pops"], b["pops"])
    assert ok, f
    assert a["cost"] == b["cost"] and np.array_equal(_bits(a["path"]), _bits(b["path"]))


def test_start_outside_grid_and_in_occupied_cell():
    """Grid3D::set_start_node (Grid3D.cpp:127-160) maps an out-of-grid start silently to cell (0, 0); an occupied start
    cell is not rejected by the reference either.  Whatever the reference returns, the device returns."""
    sc = S.c1_scenario(3)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, crm = _ctx(P), orc.ref(P)
    for o in (ctx, crm):
        S.build_map(o, sc)
    box = sc["boxes"][0]
    starts = [np.array([-500.0, 300.0, 0.1], np.float32),                 # far outside
              np.array([box[0], box[1], 0.0], np.float32)]                 # centre of an obstacle
    for s in starts:
        crm.scrub()
        a = ctx.find_path(3.0, s, max_expansions=1 << 17)
        b = crm.find_path(3.0, s)
        assert a["success"] == b["success"] and a["n_pops"] == b["n_pops"], (s, a["n_pops"], b["n_pops"])
        if b["n_pops_bin_oob"] == 0:
            ok, f = _states_equal(a["pops"], b["pops"])
            assert ok, f
            assert a["cost"] == b["cost"] and np.array_equal(_bits(a["path"]), _bits(b["path"]))


def test_small_pools_retry_gives_the_same_answer():
    """The reference's containers are unbounded; tiny device pools overflow and the query is re-run with 8x pools until it
    fits: the result must not depend on the starting capacity."""
    P = orc.ref_test_params()
    ctx = _ctx(P)
    orc.setup_ref_test_scenario(ctx)
    q = ctx.make_queries([[18.0, 18.0, np.pi / 2, 2.0]], [0])
    big = ctx.find_path_batch(q, ctx.make_opts(trace_cap=4096, path_cap=2048))
    small = ctx.find_path_batch(q, ctx.make_opts(trace_cap=4096, path_cap=2048, max_expansions=64, max_open=64, max_open2d=64))
    assert small[0][0]["status"] == 0 and big[0][0]["status"] == 0
    for f in ("success", "n_pops", "n_path", "n_chain", "n_dubins"):
        assert small[0][0][f] == big[0][0][f], f
    assert np.float32(small[0][0]["cost"]) == np.float32(big[0][0]["cost"])
    n = int(big[0][0]["n_path"])
    assert np.array_equal(_bits(small[1][0, :n]), _bits(big[1][0, :n]))
    # K-POP pools as well
    kb = ctx.find_path_batch(q, ctx.make_opts(path_cap=2048, mode=1, kpop=32))
    ks = ctx.find_path_batch(q, ctx.make_opts(path_cap=2048, mode=1, kpop=32, max_expansions=64))
    assert ks[0][0]["status"] == 0 and ks[0][0]["n_pops"] == kb[0][0]["n_pops"]
    assert np.float32(ks[0][0]["cost"]) == np.float32(kb[0][0]["cost"])


def test_path_cap_truncation_is_flagged():
    P = orc.ref_test_params()
    ctx = _ctx(P)
    orc.setup_ref_test_scenario(ctx)
    q = ctx.make_queries([[18.0, 18.0, np.pi / 2, 2.0]], [0])
    full = ctx.find_path_batch(q, ctx.make_opts(path_cap=2048))
    cut = ctx.find_path_batch(q, ctx.make_opts(path_cap=16))
    bit = _status_bits()["PP_STATUS_PATH_OVERFLOW"]
    assert full[0][0]["status"] == 0 and (cut[0][0]["status"] & bit)
    assert cut[0][0]["success"] == full[0][0]["success"] and cut[0][0]["n_pops"] == full[0][0]["n_pops"]
    assert np.float32(cut[0][0]["cost"]) == np.float32(full[0][0]["cost"])


@pytest.mark.parametrize("mode", [0, 1])
def test_batch_results_do_not_depend_on_order_or_slots(mode):
    """Queries are fetched dynamically by whichever slot is free; every query owns its scratch, so results must be the same
    for any order and any number of resident slots."""
    sc = S.c4_group(2, n_starts=24)
    P = orc.make_params(grid_size=512, resolution=0.2)
    ctx = _ctx(P)
    ctx.update_goal(sc["goal"], sc["frame_start"])
    for _ in range(sc["rounds"]):
        ctx.update_boxes(sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS)
        ctx.decay()
    cand = sc["start_candidates"][:24]
    q = ctx.make_queries(cand, [0] * len(cand))
    a = ctx.find_path_batch(q, ctx.make_opts(path_cap=1024, mode=mode, kpop=32))
    perm = np.random.RandomState(0).permutation(len(q))
    b = ctx.find_path_batch(q[perm], ctx.make_opts(path_cap=1024, mode=mode, kpop=32, max_slots=3))
    for f in ("success", "status", "n_pops", "n_path", "n_closed"):
        assert np.array_equal(a[0][f][perm], b[0][f]), f
    assert np.array_equal(_bits(a[0]["cost"][perm]), _bits(b[0]["cost"]))
    assert np.array_equal(_bits(a[1][perm]), _bits(b[1]))


def test_kpop_field_follows_map_and_goal_changes():
    """The exact 2D field is cached per group and must be rebuilt after any map or goal change."""
    sc = S.c1_scenario(1)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, port = _ctx(P), orc.port(P)
    for o in (ctx, port):
        S.build_map(o, sc)
    q0 = sc["queries"][0]
    q = ctx.make_queries([q0], [0])
    opts = ctx.make_opts(trace_cap=1 << 16, path_cap=4096, mode=1, kpop=32)

    def check():
        res, paths, curv, trace = ctx.find_path_batch(q, opts)
        h1, _, _ = ctx.field2d()
        _same(res[0], trace[0], paths[0], curv[0], port_kpop(port, float(q0[3]), q0[:3], 32, h1))
        d = orc.field2d(port)
        reach = d >= 0
        assert np.array_equal(reach, h1 < 1e30)
        assert np.max(np.abs(h1[reach] - d[reach]) / np.maximum(d[reach], 1e-9)) <= 1e-5
    check()
    extra = np.array([[12.0, 1.0, 2.0, 6.0]], np.float32); conf = np.array([0.95], np.float32)
    for o in (ctx, port):                                   # a new obstacle across the corridor
        for _ in range(3):
            o.update_boxes(extra, conf, S.APF_ADDED_RADIUS)
    check()
    new_goal = sc["goal"] + np.array([-3.0, 2.0, 0.2], np.float32)
    for o in (ctx, port):                                   # goal change: relocation + new frame
        o.update_goal(new_goal, sc["frame_start"])
    check()
