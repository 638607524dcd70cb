"""CPU tests of the ORACLES (no GPU): the restatement in oracle/port and, when present, the compiled
reference oracle/_ref are pinned against (1) the result vectors the reference's author recorded in its
plotting scripts and (2) fixtures generated from the unmodified reference (tests/golden/make_golden.py).
"""
import zlib

import numpy as np
import pytest

import orc
import scenarios as S

GOLD = orc.ROOT + "/tests/golden/"


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _flavours():
    out = [("port", orc.port)]
    if orc.have_ref():
        out += [("ref", orc.ref), ("crm", orc.crm)]
    return out


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


@pytest.mark.parametrize("name,mk", _flavours())
def test_hybrid_astar_golden_path(name, mk):
    """utils/hybrid_astar/plot.py:47-51 -- 43 poses, cost 33.0305, Dubins-shot termination."""
    o = mk(orc.ref_test_params())
    orc.setup_ref_test_scenario(o)
    r = o.find_path(2.0, orc.REF_TEST_START)
    gold = np.load(GOLD + "hybrid_astar_path.npy")
    assert r["success"] and abs(float(r["cost"]) - 33.0305) < 5e-5
    assert r["n_pops"] == 882
    assert r["path"].shape == (43, 3)
    assert np.allclose(r["path"][::-1], gold, rtol=2e-5, atol=2e-5)    # 6 significant digits in the source
    assert len(r["curvature"]) == 43 and r["curvature"][0] == 0.0


@pytest.mark.parametrize("name,mk", _flavours())
def test_dubins_golden_path(name, mk):
    """utils/dubins_paths.py:6 -- RSL, 73 samples, (0,0,0) -> (20,-20,pi/2), r_min 4.08106, step 0.5 (recorded in double)."""
    deg = np.pi / 180.0
    P = orc.make_params(step_size=0.5, wheelbase=2.269, rear_to_cg=1.1, steering=[-30 * deg, 0.0, 30 * deg],
                        curvature_weights=[0, 0, 0], num_actions=1)
    o = mk(P)
    assert abs(o.consts().r_min - 4.08106) < 1e-5
    xyh, curv, length, flag = o.dubins_path([0, 0, 0], [20, -20, np.pi / 2])
    gold = np.load(GOLD + "dubins_rsl_path.npy")
    ln, ty, _ = o.dubins_length(np.array([[0, 0, 0]], np.float32), np.array([20, -20, np.pi / 2], np.float32))
    assert ty[0] == 1                                  # RSL
    assert abs(length - 36.8303146) < 2e-5 and abs(ln[0] - length) == 0
    assert xyh.shape == gold.shape == (73, 3)
    assert np.allclose(xyh, gold, rtol=1e-4, atol=2e-4)   # float32 here vs the double run the author recorded


@pytest.mark.parametrize("name,mk", _flavours())
def test_vehicle_rollout_golden(name, mk):
    """utils/vehicle_mode.py:12 -- 33-point simulate_action sequence (utils/vehicle_dubins/test_vehicle_dubins.cpp:61-64).
    The recorded run is VehicleModel<double>; in float int(ts/dt) is 499 instead of 500, i.e. each
    step is 0.2 % short, hence the tolerance."""
    deg = np.pi / 180.0
    P = orc.make_params(step_size=0.5, max_lat_acc=4.0, max_long_dec=2.0, wheelbase=2.269, rear_to_cg=1.1, num_actions=3,
                        steering=[-30 * deg, -20 * deg, -10 * deg, 0.0, 10 * deg, 20 * deg, 30 * deg], curvature_weights=[0] * 7)
    o = mk(P)
    idx = [6] * 6 + [5] * 4 + [4] * 5 + [3] * 5 + [4] * 4 + [5] * 4 + [4] * 4
    st = np.zeros(1, orc.STATE_DT)
    st["vmin_sqr"] = 16.0; st["curvature_index"] = 3; st["angle_bin"] = 36
    pts = [[0.0, 0.0]]
    for a in idx:
        out, cnt, _ = o.rollout(st)
        succ = out[0, :cnt[0]]
        nxt = succ[succ["curvature_index"] == a]
        assert len(nxt) == 1, "primitive pruned"
        st = np.zeros(1, orc.STATE_DT)
        for f in orc.STATE_DT.names:
            st[f] = nxt[f][0]
        st["curvature_index"] = 3      # simulate_action ignores the previous action (all 7 primitives reachable)
        pts.append([float(nxt["x"][0]), float(nxt["y"][0])])
    gold = np.load(GOLD + "vehicle_rollout.npy")
    assert np.allclose(np.array(pts), gold, rtol=4e-3, atol=4e-3)


def test_port_matches_reference_fixtures():
    """The restatement against outputs of the unmodified reference (committed fixtures; no /root/reference needed)."""
    g = np.load(GOLD + "golden_search.npz")
    o = orc.port(orc.ref_test_params())
    orc.setup_ref_test_scenario(o)
    assert np.array_equal(_bits(o.get_map()), _bits(g["ref_map"]))
    r = o.find_path(2.0, orc.REF_TEST_START)
    assert np.array_equal(r["pops"], g["ref_pops"])                   # full expansion sequence incl. g, f bits
    assert np.array_equal(_bits(r["path"]), _bits(g["ref_path"])) and np.array_equal(_bits(r["curvature"]), _bits(g["ref_curv"]))
    assert r["cost"] == g["ref_cost"]


def test_port_c1_fixtures():
    rows = np.load(GOLD + "golden_c1.npz")["rows"]
    for rec in rows[:8]:
        seed = int(rec[0])
        sc = S.c1_scenario(seed)
        o = orc.port(orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"]))
        S.build_map(o, sc)
        q = sc["queries"][0]
        r = o.find_path(float(q[3]), q[:3])
        assert int(r["success"]) == int(rec[1]) and r["n_pops"] == int(rec[3]) and len(r["path"]) == int(rec[4])
        assert abs(float(r["cost"]) - rec[2]) < 1e-6
        assert zlib.crc32(r["pops"][["ci", "cj", "bin"]].tobytes()) == int(rec[5])
        assert zlib.crc32(r["path"].tobytes()) == int(rec[6])
        assert zlib.crc32(o.get_map().tobytes()) == int(rec[7])


def test_port_map_c2_fixture():
    """C2: 256 boxes into 2048^2, 3 rounds; CRC of all 4 194 304 floats per round recorded from the reference."""
    g = np.load(GOLD + "golden_map_c2.npz")
    sc = S.c2_scenario()
    o = orc.port(orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"]))
    o.update_goal(sc["goal"], sc["frame_start"])
    for k in range(sc["rounds"]):
        o.update_boxes_2d(sc["boxes"], sc["conf"])
        o.decay()
        m = o.get_map()
        assert zlib.crc32(m.tobytes()) == int(g["crcs"][k])
    assert int((m >= o.consts().log_threshold).sum()) == int(g["occupied"])


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference not present")
def test_port_vs_compiled_reference_pieces():
    """Function-by-function, bit for bit: port == oracle/_ref (same libm, same libstdc++)."""
    P = orc.ref_test_params()
    a, b = orc.port(P), orc.ref(P)
    for o in (a, b):
        orc.setup_ref_test_scenario(o)
    rs = np.random.RandomState(11)
    n = 3000
    st = np.zeros(n, orc.STATE_DT)
    st["x"] = rs.uniform(0, 30, n); st["y"] = rs.uniform(0, 30, n); st["heading"] = rs.uniform(-3.05, 3.05, n)
    st["g"] = rs.uniform(0, 50, n); st["vmin_sqr"] = rs.uniform(0, 9, n); st["f"] = st["g"]
    st["curvature_index"] = rs.randint(0, P.num_steering, n)
    prec = np.float32(b.consts().precision)
    st["angle_bin"] = ((np.round(st["heading"] / prec).astype(np.float32) * prec).astype(np.float64) + np.pi) / float(prec)
    for fn in ("rollout", "expand"):
        x, xc, xf = getattr(a, fn)(st); y, yc, yf = getattr(b, fn)(st)
        assert np.array_equal(xc, yc) and np.array_equal(xf, yf) and x.tobytes() == y.tobytes(), fn
    xyh = np.stack([st["x"], st["y"], st["heading"]], 1)
    assert np.array_equal(_bits(a.apf(xyh)), _bits(b.apf(xyh)))
    goal = np.array(list(b.consts().goal_grid), np.float32)
    for u, v in zip(a.dubins_length(xyh, goal), b.dubins_length(xyh, goal)):
        assert u.tobytes() == v.tobytes()
    free = np.argwhere(b.get_map() < b.consts().log_threshold)
    ij = free[rs.choice(len(free), 400, replace=False)].astype(np.int32)
    a.scrub(); b.scrub()
    assert np.array_equal(_bits(a.astar_lazy(ij)), _bits(b.astar_lazy(ij)))
    va, ga, fa = a.astar_dump(); vb, gb, fb = b.astar_dump()
    assert np.array_equal(va, vb) and np.array_equal(_bits(fa), _bits(fb))


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference not present")
def test_port_vs_compiled_reference_search_c4():
    sc = S.c4_group(3, n_starts=3)
    P = orc.make_params(grid_size=512, resolution=0.2)
    a, b = orc.port(P), orc.ref(P)
    for o in (a, b):
        S.build_map(o, sc)
    assert np.array_equal(_bits(a.get_map()), _bits(b.get_map()))
    compared = 0
    for q in S.select_starts(sc, b.get_map(), b.consts().log_threshold, b.set_start):
        a.scrub(); b.scrub()
        x = a.find_path(float(q[3]), q[:3]); y = b.find_path(float(q[3]), q[:3])
        if y["n_pops_bin_oob"] > 0:
            continue   # the reference reads _offset_xy[.][72] out of bounds there; undefined, not compared
        compared += 1
        assert x["n_pops"] == y["n_pops"] and np.array_equal(x["pops"], y["pops"])
        assert x["cost"] == y["cost"] and np.array_equal(_bits(x["path"]), _bits(y["path"]))
    assert compared >= 1


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference not present")
def test_relocation_port_vs_reference():
    """Grid3D::relocate_obstacles on a non-empty map (goal change), port == reference bit for bit."""
    P = orc.make_params(grid_size=120, resolution=0.3)
    a, b = orc.port(P), orc.ref(P)
    boxes = np.array([[8, 2, 2, 3], [14, -3, 1.5, 1.5], [20, 6, 4, 1]], np.float32)
    for o in (a, b):
        o.update_goal([20, 5, 0.2], [0, 0, 0])
        o.update_boxes(boxes, np.full(3, 0.9, np.float32), 1.5)
        o.update_goal([22, 9, -0.1], [1.5, 0.4, 0.1])
    assert (b.get_map() != 0).sum() > 50
    assert np.array_equal(_bits(a.get_map()), _bits(b.get_map()))


def test_fp32_dubins_restatement_close_to_reference():
    """oracle/port/fmath.inc (libm mode 2, the K-POP heuristic flavour): Dubins lengths within 1e-5 relative of the
    unmodified reference, except at the +-2pi branch flips of the reference formula (counted)."""
    P = orc.ref_test_params()
    ref, port = orc.ref(P), orc.port(P)
    for o in (ref, port):
        orc.setup_ref_test_scenario(o)
    rs = np.random.RandomState(5)
    n = 100000
    goal = np.array(list(ref.consts().goal_grid), np.float32)
    starts = np.stack([rs.uniform(0, 30, n), rs.uniform(0, 30, n), rs.uniform(-3.14, 3.14, n)], 1).astype(np.float32)
    r, _, _ = ref.dubins_length(starts, goal)
    port.lib.port_set_libm(2)
    try:
        a, _, _ = port.dubins_length(starts, goal)
    finally:
        port.lib.port_set_libm(0)
    rel = np.abs(a - r) / np.maximum(np.abs(r), 1e-6)
    jumps = int((rel > 1e-5).sum())
    assert jumps <= n // 2000, jumps
    assert rel[rel <= 1e-5].max() < 1e-5
    print(f"fp32 restatement vs reference: max rel {rel[rel <= 1e-5].max():.3g}, flips {jumps}/{n}")


@pytest.mark.skipif(not orc.have_ref(), reason="compiled reference not present")
@pytest.mark.parametrize("seed", [0, 3])
def test_port_session_on_one_object_equals_reference(seed):
    """the restatement carries the 2D heuristic cache from query to query like one reference object does (11-query
    session incl. a bare reset and a second waypoint whose update_goal relocates the non-empty map)."""
    import scenarios as S
    sc, ops = S.session_ops(seed, goal_changes=True)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    a, b = orc.port(P), orc.ref(P)
    ra, rb = S.run_session(a, ops), S.run_session(b, ops)
    assert len(ra) == len(rb) == 11
    for k, (x, y) in enumerate(zip(ra, rb)):
        if y["n_pops_bin_oob"]:
            continue
        assert x["n_pops"] == y["n_pops"] and np.array_equal(x["pops"], y["pops"]), k
        assert x["cost"] == y["cost"] and np.array_equal(x["path"].view(np.uint32), y["path"].view(np.uint32)), k
