"""Synthetic scenario generators for BASELINE.json's configs.

All random inputs are generated once on the host from a fixed seed and handed, as the same float32
arrays, to the oracle and to the GPU path.  Shared by tests/ and bench.py (no oracle imports here).
"""
import numpy as np

APF_ADDED_RADIUS = 1.5   # launch/local_planner.launch apf_object_added_radius


def _u(rs, n):
    """uniform [0,1) with 24 random bits (exactly representable in float32)."""
    return (rs.randint(0, 1 << 24, size=n).astype(np.float64)) * (2.0 ** -24)


def launch_params_kwargs(grid_size, resolution):
    """Launch defaults (launch/local_planner.launch:11-45) at a given grid shape."""
    return dict(grid_size=grid_size, resolution=resolution)


def c1_scenario(seed):
    """C1: N=200, res 0.2, 5 boxes, one query from the frame origin."""
    rs = np.random.RandomState(seed)
    side = 1.0 + 2.0 * _u(rs, 5)
    cx = 6.0 + 14.0 * _u(rs, 5)
    cy = -6.0 + 14.0 * _u(rs, 5)
    boxes = np.stack([cx, cy, side, side], 1).astype(np.float32)
    return dict(grid_size=200, resolution=0.2, goal=np.array([24.0, 4.0, 0.3], np.float32),
                frame_start=np.zeros(3, np.float32), boxes=boxes, conf=np.full(5, 0.85, np.float32),
                rounds=4, queries=np.array([[0.0, 0.0, 0.0, 3.0]], np.float32))


def c4_group(seed, n_boxes=96, n_starts=64, grid_size=512, resolution=0.2):
    """One (map, goal) group of C4/C5: clutter corridor + `n_starts` start poses."""
    rs = np.random.RandomState(1000 + seed)
    side = 1.0 + 2.0 * _u(rs, n_boxes)
    cx = 15.0 + 36.0 * _u(rs, n_boxes)
    cy = -15.0 + 35.0 * _u(rs, n_boxes)
    boxes = np.stack([cx, cy, side, side], 1).astype(np.float32)
    # over-generate starts; occupied ones are rejected by the caller once the map exists
    m = 4 * n_starts
    sx = -5.0 + 15.0 * _u(rs, m)
    sy = -10.0 + 20.0 * _u(rs, m)
    sh = -0.6 + 1.2 * _u(rs, m)
    sv = 5.0 * _u(rs, m)
    cand = np.stack([sx, sy, sh, sv], 1).astype(np.float32)
    L = grid_size * resolution
    goal = np.array([0.6 * L, 0.1 * L, 0.3], np.float32)   # (61.44, 10.24, 0.3) at 512 x 0.2
    return dict(grid_size=grid_size, resolution=resolution, goal=goal, frame_start=np.zeros(3, np.float32),
                boxes=boxes, conf=np.full(n_boxes, 0.85, np.float32), rounds=4, start_candidates=cand,
                n_starts=n_starts)


def build_map(planner, sc):
    """update_goal + `rounds` x (boxes, decay) on any object with the oracle-style interface."""
    planner.update_goal(sc["goal"], sc["frame_start"])
    planner.reset()
    for _ in range(sc["rounds"]):
        planner.update_boxes(sc["boxes"], sc["conf"], APF_ADDED_RADIUS)
        planner.decay()


def select_starts(sc, grid_map, log_threshold, set_start_fn):
    """Keep the first n_starts candidates whose start cell is free (reject occupied starts)."""
    out = []
    for q in sc["start_candidates"]:
        st = set_start_fn(q[:3])
        if grid_map[st["ci"], st["cj"]] < log_threshold:
            out.append(q)
        if len(out) == sc["n_starts"]:
            break
    return np.array(out, np.float32)


def c2_scenario(seed=42, n_boxes=256, grid_size=2048, resolution=0.2):
    """C2: 256 boxes rasterised into a 2048^2 log-odds map, 3 rounds of (boxes, decay)."""
    rs = np.random.RandomState(seed)
    L = grid_size * resolution
    goal = np.array([0.3 * L, 0.1 * L, 0.0], np.float32)      # (122.88, 40.96, 0)
    dx = 0.5 + 5.5 * _u(rs, n_boxes)
    dy = 0.5 + 5.5 * _u(rs, n_boxes)
    # centres uniform over the map extent, expressed in the grid frame then rotated back to the world
    gx = L * _u(rs, n_boxes) - 0.8 * L
    gy = L * _u(rs, n_boxes) - 0.5 * L
    h = np.arctan2(goal[1], goal[0])
    wx = gx * np.cos(h) - gy * np.sin(h) + goal[0]
    wy = gx * np.sin(h) + gy * np.cos(h) + goal[1]
    boxes = np.stack([wx, wy, dx, dy], 1).astype(np.float32)
    conf = (0.55 + 0.44 * _u(rs, n_boxes)).astype(np.float32)
    return dict(grid_size=grid_size, resolution=resolution, goal=goal, frame_start=np.zeros(3, np.float32),
                boxes=boxes, conf=conf, rounds=3)


def session_ops(seed, goal_changes=True, n_ticks=6):
    """A planner SESSION on one object, as src/local_planner.cpp drives it: waypoint -> update_goal + reset,
    then per tick new detections (boxes + decay) and one find_path from the advancing vehicle pose, with NO reset in between;
    later a bare reset, and (optionally) a second waypoint whose update_goal relocates the non-empty map.  C1 shape
    (N = 200, 0.2 m).  Returns a list of ops: ("goal", goal3, start3) | ("reset",) | ("boxes", boxes, conf) | ("decay",) |
    ("query", vel, start3)."""
    rs = np.random.RandomState(7000 + seed)
    sc = c1_scenario(seed)
    ops = [("goal", sc["goal"], sc["frame_start"]), ("reset",)]
    for _ in range(sc["rounds"]):
        ops += [("boxes", sc["boxes"], sc["conf"]), ("decay",)]
    pose = np.array([0.0, 0.0, 0.0], np.float32)

    def tick(pose, k):
        extra = np.array([[8.0 + 10.0 * _u(rs, 1)[0], -4.0 + 10.0 * _u(rs, 1)[0], 1.0 + _u(rs, 1)[0], 1.0 + _u(rs, 1)[0]]], np.float32)
        boxes = np.concatenate([sc["boxes"], extra]).astype(np.float32)
        return [("boxes", boxes, np.full(len(boxes), 0.85, np.float32)), ("decay",),
                ("query", np.float32(2.0 + 0.2 * k), pose.copy())]

    for k in range(n_ticks):
        ops += tick(pose, k)
        pose = (pose + np.array([0.9, 0.12 * (k % 3 - 1), 0.03 * (k % 2)], np.float32)).astype(np.float32)
    ops += [("reset",)] + tick(pose, n_ticks) + tick(pose, n_ticks + 1)
    if goal_changes:
        goal2 = np.array([30.0, -6.0, -0.2], np.float32)
        ops += [("goal", goal2, pose.copy()), ("reset",)]
        for k in range(3):
            ops += tick(pose, k)
            pose = (pose + np.array([0.8, -0.2, -0.02], np.float32)).astype(np.float32)
    return sc, ops


def run_session(planner, ops, fresh_each_query=None):
    """Apply `ops` to any object with the oracle-style interface; returns the find_path results in order.
    `fresh_each_query(planner)` is called before every query when given (the fresh-cache comparison)."""
    out = []
    for op in ops:
        if op[0] == "goal":
            planner.update_goal(op[1], op[2])
        elif op[0] == "reset":
            planner.reset()
        elif op[0] == "boxes":
            planner.update_boxes(op[1], op[2], APF_ADDED_RADIUS)
        elif op[0] == "decay":
            planner.decay()
        elif op[0] == "query":
            if fresh_each_query:
                fresh_each_query(planner)
            out.append(planner.find_path(float(op[1]), op[2]))
    return out
