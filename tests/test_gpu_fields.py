"""GPU parity tests of the heuristic-field kernels (north_star (d), (e); BASELINE config C3) and of the map kernels at
the C2 shape.  Tolerances are the ones north_star states: Dubins 1e-5 relative; the 2D field against a double-precision
Dijkstra 1e-5 relative (it cannot equal AStar::find_path, SURVEY F4 -- the distribution of that difference is printed)."""
import zlib

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


def test_map_c2_shape_bitexact():
    """C2: 256 boxes into a 2048^2 log-odds map, 3 rounds of (boxes, decay): CRC of all 4 194 304 floats per round
    against the fixture recorded from the unmodified reference; and bit for bit against the port oracle."""
    g = np.load(orc.ROOT + "/tests/golden/golden_map_c2.npz")
    sc = S.c2_scenario()
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, port = _ctx(P), orc.port(P)
    for o in (ctx, port):
        o.update_goal(sc["goal"], sc["frame_start"])
    for k in range(sc["rounds"]):
        for o in (ctx, port):
            o.update_boxes_2d(sc["boxes"], sc["conf"])
            o.decay()
        m = ctx.get_map()
        assert zlib.crc32(m.tobytes()) == int(g["crcs"][k])
        assert np.array_equal(_bits(m), _bits(port.get_map()))
    assert int((m >= ctx.consts().log_threshold).sum()) == int(g["occupied"])


def test_degenerate_lane_lines_bitexact():
    """Zero-length / NaN / overflowing lane lines: the reference draws nothing for them (x86 float -> int gives INT_MIN, which
    fails its bounds test, Grid2D.cpp:163-172); the goal cell in particular must stay free."""
    P = orc.make_params(grid_size=64, resolution=0.5)
    ctx, ref = _ctx(P), orc.ref(P)
    lines = np.array([[3, 1, 3, 1], [np.nan, 0, 4, 4], [2, -3, 9, 5], [1e30, 0, -1e30, 0]], np.float32)
    for x in (ctx, ref):
        x.update_goal([10, 0, 0], [0, 0, 0])
        x.update_lines(lines, np.full(len(lines), 0.8, np.float32), 1.0)
    a, b = ctx.get_map(), ref.get_map()
    assert np.array_equal(_bits(a), _bits(b))
    fr = ctx.frame()
    assert a[fr.goal_ci, fr.goal_cj] == 0.0 and (a != 0).sum() > 10


def test_relocation_bitexact():
    """Grid3D::relocate_obstacles (goal change on a non-empty map) on the device == port == reference."""
    P = orc.make_params(grid_size=120, resolution=0.3)
    ctx, port = _ctx(P), orc.port(P)
    boxes = np.array([[8, 2, 2, 3], [14, -3, 1.5, 1.5], [20, 6, 4, 1]], np.float32)
    for o in (ctx, port):
        o.update_goal([20, 5, 0.2], [0, 0, 0])
        o.update_boxes(boxes, np.full(3, 0.9, np.float32), 1.5)
        o.update_goal([22, 9, -0.1], [1.5, 0.4, 0.1])
        o.decay()
    assert (port.get_map() > 0).sum() > 50
    assert np.array_equal(_bits(ctx.get_map()), _bits(port.get_map()))
    if orc.have_ref():
        ref = orc.ref(P)
        ref.update_goal([20, 5, 0.2], [0, 0, 0]); ref.update_boxes(boxes, np.full(3, 0.9, np.float32), 1.5)
        ref.update_goal([22, 9, -0.1], [1.5, 0.4, 0.1]); ref.decay()
        assert np.array_equal(_bits(ctx.get_map()), _bits(ref.get_map()))


@pytest.mark.parametrize("seed,n", [(0, 512), (3, 512), (1, 200)])
def test_field2d_vs_dijkstra(seed, n):
    sc = S.c4_group(seed, n_starts=1, grid_size=n, resolution=0.2 * 512 / n if n != 200 else 0.2)
    P = orc.make_params(grid_size=n, resolution=sc["resolution"])
    ctx, port = _ctx(P), orc.port(P)
    for o in (ctx, port):
        S.build_map(o, sc)
    f, sweeps, ms = ctx.field2d()
    d = orc.field2d(port)
    reach = d >= 0
    assert np.array_equal(reach, f < 1e30), "reachability differs"
    rel = np.abs(f[reach].astype(np.float64) - d[reach]) / np.maximum(d[reach], 1e-9)
    print(f"field2d N={n}: {sweeps} sweeps, {ms:.3f} ms, {reach.mean() * 100:.1f}% reachable, max rel err {rel.max():.3g}")
    assert rel.max() <= 1e-5
    # how far the reference's lazy cached A* is from the true field (reported)
    free = np.argwhere(reach & (d > 0))
    rs = np.random.RandomState(0)
    ij = free[rs.choice(len(free), min(300, len(free)), replace=False)].astype(np.int32)
    port.scrub()
    lazy = port.astar_lazy(ij).astype(np.float64)
    true = d[ij[:, 0], ij[:, 1]]
    print(f"  reference lazy A* vs true distance on {len(ij)} cells: equal {np.mean(np.abs(lazy - true) <= 1e-4 * true) * 100:.0f}%, "
          f"max overestimate {np.max(lazy / true - 1) * 100:.1f}%")
    assert (lazy >= true * (1 - 1e-5)).all()


def test_dubins_field_vs_reference():
    """N = 96: every (i, j, bin) state against Dubins<float>::get_shortest_path_length (port oracle, stock libm)."""
    n = 96
    P = orc.make_params(grid_size=n, resolution=0.4)
    ctx, port = _ctx(P), orc.port(P)
    for o in (ctx, port):
        o.update_goal([25, 3, 0.3], [0, 0, 0])
    f, ms = ctx.field3d(use_h2d=False)
    c = port.consts()
    goal = np.array(list(c.goal_grid), np.float32)
    prec = np.float32(c.precision)
    ii, jj, bb = np.meshgrid(np.arange(n), np.arange(n), np.arange(72), indexing="ij")
    head = (-np.pi + (bb.astype(np.float32) * prec).astype(np.float64)).astype(np.float32)
    starts = np.stack([(ii * np.float32(0.4)).astype(np.float32), (jj * np.float32(0.4)).astype(np.float32), head], -1).reshape(-1, 3)
    ref_len, _, _ = port.dubins_length(starts, goal)
    got = f.reshape(-1)
    rel = np.abs(got - ref_len) / np.maximum(ref_len, 1e-6)
    flips = int((rel > 1e-5).sum())
    print(f"dubins field: {len(got)} states in {ms:.3f} ms, max rel err (non-flip) {rel[rel <= 1e-5].max():.3g}, +-2pi branch flips {flips}")
    assert flips <= len(got) // 1000          # ulp-level flips of the +-2pi corrections (discontinuities of the formula)
    # with the 2D field folded in
    S.build_map(ctx, dict(goal=np.array([25, 3, 0.3], np.float32), frame_start=np.zeros(3, np.float32),
                          boxes=np.array([[10, 2, 3, 3]], np.float32), conf=np.array([0.9], np.float32), rounds=2))
    h2d, _, _ = ctx.field2d()
    f2, _ = ctx.field3d(use_h2d=True)
    f1, _ = ctx.field3d(use_h2d=False)
    assert np.array_equal(_bits(f2), _bits(np.maximum(h2d[:, :, None], f1)))


def test_dubins_field_c3_full_size_sample():
    """BASELINE configs[2] at its full size, 2048 x 2048 x 72 = 302 M states: every 64th state (stride 61 over the flat
    index, so that all heading bins and both grid axes are visited) against Dubins<float>::get_shortest_path_length of the
    reference arithmetic (port oracle, stock libm).  Bar: 1e-5 relative (north_star), +-2pi branch flips counted."""
    sc = S.c2_scenario()
    n = sc["grid_size"]
    P = orc.make_params(grid_size=n, resolution=sc["resolution"])
    ctx, port = _ctx(P), orc.port(P)
    for o in (ctx, port):
        o.update_goal(sc["goal"], sc["frame_start"])
    f, ms = ctx.field3d(use_h2d=False)
    c = port.consts()
    goal = np.array(list(c.goal_grid), np.float32)
    prec, res = np.float32(c.precision), np.float32(sc["resolution"])
    flat = np.arange(0, n * n * 72, 61, dtype=np.int64)
    ii, jj, bb = flat // (n * 72), (flat // 72) % n, flat % 72
    head = (-np.pi + (bb.astype(np.float32) * prec).astype(np.float64)).astype(np.float32)
    starts = np.stack([(ii.astype(np.float32) * res).astype(np.float32), (jj.astype(np.float32) * res).astype(np.float32), head], -1)
    ref_len, _, _ = port.dubins_length(starts, goal)
    got = f.reshape(-1)[flat]
    rel = np.abs(got - ref_len) / np.maximum(ref_len, 1e-6)
    flips = int((rel > 1e-5).sum())
    print(f"C3 dubins field: {n}x{n}x72 in {ms:.3f} ms; {len(flat)} sampled states, max rel err (non-flip) "
          f"{rel[rel <= 1e-5].max():.3g}, +-2pi branch flips {flips}")
    assert flips <= len(flat) // 1000
