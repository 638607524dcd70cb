"""ctypes access to the CPU oracles (TEST INFRASTRUCTURE ONLY).

Loads oracle/_ref/libref_oracle.so ("ref": unmodified reference, stock glibc float libm),
oracle/_ref/libref_oracle_crm.so ("crm": same objects, pinned correctly-rounded float
transcendentals in the device-executed functions, see oracle/cr_math.c) and
oracle/libpp_oracle_port.so ("port": the CPU restatement in oracle/port/).
Nothing in the product package imports this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_STEER = 16


class Params(C.Structure):
    _fields_ = [
        ("shot_interval", C.c_int), ("shot_decay", C.c_int),
        ("resolution", C.c_float), ("obstacle_threshold", C.c_float),
        ("prob_min", C.c_float), ("prob_max", C.c_float), ("prob_free", C.c_float),
        ("grid_size", C.c_int), ("allow_diag", C.c_int),
        ("step_size", C.c_float), ("max_lat_acc", C.c_float), ("max_long_dec", C.c_float),
        ("wheelbase", C.c_float), ("rear_to_cg", C.c_float),
        ("apf_rep_constant", C.c_float), ("apf_active_angle", C.c_float),
        ("num_angle_bins", C.c_int), ("num_actions", C.c_int), ("num_steering", C.c_int),
        ("steering", C.c_float * MAX_STEER), ("curvature_weights", C.c_float * MAX_STEER),
    ]


STATE_DT = np.dtype([("x", "f4"), ("y", "f4"), ("heading", "f4"), ("g", "f4"), ("f", "f4"),
                     ("vmin_sqr", "f4"), ("curvature_index", "i4"), ("angle_bin", "i4"),
                     ("ci", "i4"), ("cj", "i4")])
POP_DT = np.dtype([("ci", "i4"), ("cj", "i4"), ("bin", "i4"), ("x", "f4"), ("y", "f4"),
                   ("heading", "f4"), ("g", "f4"), ("f", "f4")])


class Consts(C.Structure):
    _fields_ = [("log_threshold", C.c_float), ("log_min", C.c_float), ("log_max", C.c_float),
                ("log_free", C.c_float), ("grid_heading", C.c_float),
                ("goal_world", C.c_float * 3), ("goal_grid", C.c_float * 3),
                ("goal_bin", C.c_int), ("goal_ci", C.c_int), ("goal_cj", C.c_int),
                ("precision", C.c_float), ("r_min", C.c_float), ("ang_step", C.c_float),
                ("num_apf", C.c_int)]


class Result(C.Structure):
    _fields_ = [("success", C.c_int), ("cost", C.c_float), ("n_path", C.c_int),
                ("n_pops", C.c_int), ("n_pops_bin_oob", C.c_int)]


def make_params(**kw):
    """Launch defaults of the reference (launch/local_planner.launch:11-45)."""
    d = dict(shot_interval=100, shot_decay=10, resolution=0.3, obstacle_threshold=0.7,
             prob_min=0.05, prob_max=0.975, prob_free=0.45, grid_size=100, allow_diag=1,
             step_size=0.4, max_lat_acc=2.0, max_long_dec=2.5, wheelbase=2.269, rear_to_cg=1.135,
             apf_rep_constant=1.0, apf_active_angle=float(np.float32(180.0 * (np.pi / 180.0))),
             num_angle_bins=72, num_actions=2,
             steering=[float(np.float32(a) * np.float32(np.pi / 180.0)) for a in (-40, -20, 0, 20, 40)],
             curvature_weights=[1.0, 0.5, 0.0, 0.5, 1.0])
    d.update(kw)
    p = Params()
    for k, v in d.items():
        if k in ("steering", "curvature_weights"):
            continue
        setattr(p, k, v)
    st = d["steering"]
    cw = list(d["curvature_weights"]) + [0.0] * MAX_STEER
    p.num_steering = len(st)
    for i, a in enumerate(st):
        p.steering[i] = a
        p.curvature_weights[i] = cw[i]
    return p


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One planner object behind the oracle ABI (oracle/oracle_api.h)."""

    def __init__(self, lib, prefix, params):
        self.lib, self.pf, self.params = lib, prefix, params
        self._fn("create").restype = C.c_void_p
        self.h = C.c_void_p(self._fn("create")(C.byref(params)))
        self.N = params.grid_size
        self.stride = 2 * params.num_actions + 1

    def _fn(self, name):
        return getattr(self.lib, f"{self.pf}_{name}")

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- map / frame -----------------------------------------------------------------------
    def update_goal(self, goal, start):
        g = np.asarray(goal, np.float32); s = np.asarray(start, np.float32)
        self._fn("update_goal")(self.h, _fp(g), _fp(s))

    def reset(self):
        self._fn("reset")(self.h)

    def scrub(self):
        self._fn("scrub")(self.h)

    def update_boxes(self, boxes, conf, apf_added_radius):
        b = np.ascontiguousarray(boxes, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._fn("update_boxes")(self.h, _fp(b), _fp(c), C.c_int(len(c)), C.c_float(apf_added_radius))

    def update_boxes_2d(self, boxes, conf):
        b = np.ascontiguousarray(boxes, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._fn("update_boxes_2d")(self.h, _fp(b), _fp(c), C.c_int(len(c)))

    def update_lines(self, lines, conf, width):
        b = np.ascontiguousarray(lines, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._fn("update_lines")(self.h, _fp(b), _fp(c), C.c_int(len(c)), C.c_float(width))

    def decay(self):
        self._fn("decay")(self.h)

    def get_map(self):
        out = np.empty((self.N, self.N), np.float32)
        self._fn("get_map")(self.h, _fp(out))
        return out

    def set_map(self, m):
        m = np.ascontiguousarray(m, np.float32)
        assert m.shape == (self.N, self.N)
        self._fn("set_map")(self.h, _fp(m))

    def consts(self):
        c = Consts()
        self._fn("get_consts")(self.h, C.byref(c))
        return c

    def apf_list(self):
        n = self.consts().num_apf
        out = np.empty((n, 3), np.float32)
        self._fn("get_apf")(self.h, _fp(out))
        return out

    def tables(self):
        S, B = self.params.num_steering, self.params.num_angle_bins
        oxy = np.empty((S, B, 2), np.float32); oh = np.empty(S, np.float32)
        ac = np.empty(S, np.float32); cu = np.empty(S, np.float32)
        self._fn("get_tables")(self.h, _fp(oxy), _fp(oh), _fp(ac), _fp(cu))
        return oxy, oh, ac, cu

    # -- stateless pieces --------------------------------------------------------------------
    def set_start(self, start):
        s = np.asarray(start, np.float32)
        out = np.zeros(1, STATE_DT)
        self._fn("set_start")(self.h, _fp(s), _fp(out))
        return out[0]

    def _succ(self, name, states):
        states = np.ascontiguousarray(states, STATE_DT)
        n = len(states)
        out = np.zeros((n, self.stride), STATE_DT)
        cnt = np.zeros(n, np.int32); fl = np.zeros(n, np.int32)
        self._fn(name)(self.h, _fp(states), C.c_int(n), _fp(out), _fp(cnt), _fp(fl))
        return out, cnt, fl

    def rollout(self, states):
        return self._succ("rollout_batch", states)

    def expand(self, states):
        return self._succ("expand_batch", states)

    def apf(self, xyh):
        xyh = np.ascontiguousarray(xyh, np.float32)
        out = np.empty(len(xyh), np.float32)
        self._fn("apf_batch")(self.h, _fp(xyh), C.c_int(len(xyh)), _fp(out))
        return out

    def check_path(self, xyh):
        xyh = np.ascontiguousarray(xyh, np.float32)
        self._fn("check_path").restype = C.c_int
        return bool(self._fn("check_path")(self.h, _fp(xyh), C.c_int(len(xyh))))

    def dubins_length(self, starts, goal):
        starts = np.ascontiguousarray(starts, np.float32); goal = np.asarray(goal, np.float32)
        n = len(starts)
        ln = np.empty(n, np.float32); ty = np.empty(n, np.int32); pr = np.empty((n, 4), np.float32)
        self._fn("dubins_length_batch")(self.h, _fp(starts), C.c_int(n), _fp(goal), _fp(ln), _fp(ty), _fp(pr))
        return ln, ty, pr

    def dubins_path(self, start, goal, cap=4096):
        s = np.asarray(start, np.float32); g = np.asarray(goal, np.float32)
        xyh = np.empty((cap, 3), np.float32); cv = np.empty(cap, np.float32)
        ln = C.c_float(); fl = C.c_int()
        f = self._fn("dubins_path"); f.restype = C.c_int
        n = f(self.h, _fp(s), _fp(g), _fp(xyh), _fp(cv), C.c_int(cap), C.byref(ln), C.byref(fl))
        assert n <= cap
        return xyh[:n].copy(), cv[:n].copy(), ln.value, bool(fl.value)

    def astar_lazy(self, ij):
        ij = np.ascontiguousarray(ij, np.int32)
        out = np.empty(len(ij), np.float32)
        self._fn("astar_lazy_batch")(self.h, _fp(ij), C.c_int(len(ij)), _fp(out))
        return out

    def astar_dump(self):
        N = self.N
        v = np.empty((N, N), np.uint8); g = np.empty((N, N), np.float32); f = np.empty((N, N), np.float32)
        self._fn("astar_dump")(self.h, _fp(v), _fp(g), _fp(f))
        return v, g, f

    # -- generic footprint collision check (port / emu only: new semantics, oracle/port/footprint.inc) --
    def footprint_table(self, bin_, length, width, rear, cap=4096):
        offs = np.zeros((cap, 2), np.int16)
        fn = self._fn("footprint_table"); fn.restype = C.c_int
        n = fn(self.h, C.c_int(bin_), C.c_float(length), C.c_float(width), C.c_float(rear), _fp(offs), C.c_int(cap))
        return offs[:n].copy()

    def footprint_check(self, xyh, length, width, rear):
        p = np.ascontiguousarray(xyh, np.float32); n = len(p)
        free = np.zeros(n, np.int32); cells = np.zeros((n, 2), np.int32); hits = np.zeros(n, np.int32)
        self._fn("footprint_check")(self.h, _fp(p), C.c_int(n), C.c_float(length), C.c_float(width), C.c_float(rear),
                                    _fp(free), _fp(cells), _fp(hits))
        return free, cells, hits

    # -- velocity profile / trajectory --------------------------------------------
    def velocity_profile(self, lim5, vel_init, vcap, path_xyh, curvature, coast=False, stop=False):
        """ref only: VelocityGenerator<float>::generate_velocity_profile on a path in find_path's order (goal -> start)."""
        p = np.ascontiguousarray(path_xyh, np.float32); cv = np.ascontiguousarray(curvature, np.float32)
        vel = np.zeros(len(cv), np.float32); lim = np.asarray(lim5, np.float32)
        fn = self._fn("velocity_profile"); fn.restype = C.c_int
        ok = fn(_fp(lim), C.c_float(vel_init), C.c_float(vcap), _fp(p), _fp(cv), C.c_int(len(cv)), C.c_int(int(coast)), C.c_int(int(stop)), _fp(vel))
        return vel, bool(ok)

    def trajectory(self, lim5, vcap, stop):
        """emu only: the trajectory of the last find_path (= pp_trajectory_batch for that query): (4, m) array, feasible."""
        traj = np.zeros(4 * 4096, np.float32); ok = C.c_int(); lim = np.asarray(lim5, np.float32)
        fn = self._fn("trajectory"); fn.restype = C.c_int
        m = fn(self.h, _fp(lim), C.c_float(vcap), C.c_int(int(stop)), _fp(traj), C.byref(ok))
        return traj[:4 * m].reshape(4, m).copy(), bool(ok.value)

    # -- the search --------------------------------------------------------------------------
    def find_path(self, vel, start, pop_cap=1 << 20, path_cap=1 << 14):
        s = np.asarray(start, np.float32)
        res = Result()
        path = np.empty((path_cap, 3), np.float32); cv = np.empty(path_cap, np.float32)
        pops = np.zeros(pop_cap, POP_DT)
        self._fn("find_path")(self.h, C.c_float(vel), _fp(s), C.byref(res), _fp(path), _fp(cv),
                              C.c_int(path_cap), _fp(pops), C.c_int(pop_cap))
        n = min(res.n_path, path_cap)
        return dict(success=bool(res.success), cost=np.float32(res.cost), path=path[:n].copy(),
                    curvature=cv[:n].copy(), pops=pops[:min(res.n_pops, pop_cap)].copy(),
                    n_pops=res.n_pops, n_pops_bin_oob=res.n_pops_bin_oob)


_LIBS = {}


def _load(path):
    if path not in _LIBS:
        _LIBS[path] = C.CDLL(path)
    return _LIBS[path]


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
CRM_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle_crm.so")
PORT_SO = os.path.join(ROOT, "oracle", "libpp_oracle_port.so")


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(CRM_SO)


def have_port():
    return os.path.exists(PORT_SO)


def ref(params):
    return Oracle(_load(REF_SO), "ref", params)


def crm(params):
    return Oracle(_load(CRM_SO), "ref", params)


def port(params):
    return Oracle(_load(PORT_SO), "port", params)


_HASH_P = np.uint64(1099511628211)


def path_hash(path_xyh, curvature):
    """Same polynomial hash as oracle/ref_driver.cpp::path_hash over one query's path (n x 3 float32) + curvature (n)."""
    w = np.concatenate([np.ascontiguousarray(path_xyh, np.float32).reshape(-1).view(np.uint32),
                        np.ascontiguousarray(curvature, np.float32).reshape(-1).view(np.uint32)]).astype(np.uint64)
    if len(w) == 0:
        return 0
    with np.errstate(over="ignore"):
        pw = np.cumprod(np.full(len(w), _HASH_P, np.uint64))
        return int(np.sum((w + np.uint64(1)) * pw, dtype=np.uint64))


def ref_batch(P, groups, queries, qgroups, maps, idx=None, n_threads=None, so=None):
    """The unmodified reference on n_threads host threads (one HybridAStar<float> per thread, scrubbed per query) over
    queries[idx]: `groups` = scenario dicts (goal, frame_start, boxes, conf), `maps` = their final maps.  Returns a dict of
    per-query arrays (cost, success, pops, pops_oob, n_path, hash, busy_s) and the wall-clock seconds of the batch."""
    lib = _load(so or REF_SO)
    lib.ref_bench_queries_ex.restype = C.c_double
    G = len(groups)
    N = int(P.grid_size)
    frames = np.zeros((G, 6), np.float32)
    nb = len(groups[0]["boxes"])
    boxes = np.zeros((G, nb, 4), np.float32); conf = np.zeros((G, nb), np.float32)
    for g, sc in enumerate(groups):
        frames[g, :3] = sc["goal"]; frames[g, 3:] = sc["frame_start"]
        boxes[g] = sc["boxes"]; conf[g] = sc["conf"]
    mp = np.ascontiguousarray(np.stack(maps), np.float32)
    assert mp.shape == (G, N, N)
    if idx is None:
        idx = np.arange(len(queries))
    q4 = np.ascontiguousarray(np.asarray(queries)[idx], np.float32)
    go = np.ascontiguousarray(np.asarray(qgroups)[idx], np.int32)
    n = len(q4)
    out = dict(cost=np.zeros(n, np.float32), success=np.zeros(n, np.int32), pops=np.zeros(n, np.int32),
               pops_oob=np.zeros(n, np.int32), n_path=np.zeros(n, np.int32), hash=np.zeros(n, np.uint64),
               busy_s=np.zeros(n, np.float64))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    secs = lib.ref_bench_queries_ex(C.byref(P), vp(frames), vp(mp), vp(boxes), vp(conf), C.c_int(nb), C.c_float(1.5), C.c_int(G),
                                    vp(q4), vp(go), C.c_int(n), C.c_int(n_threads or (os.cpu_count() or 1)),
                                    vp(out["cost"]), vp(out["success"]), vp(out["pops"]), vp(out["pops_oob"]), vp(out["n_path"]),
                                    vp(out["hash"]), vp(out["busy_s"]))
    out["secs"] = float(secs)
    return out


# ---- the reference's own scenario (utils/hybrid_astar/test_hybrid_astar.cpp:13-91) ---------------
def ref_test_params():
    st = [float(np.float32(a) * np.float32(np.pi) / np.float32(180.0)) for a in (-30, -15, 0, 15, 30)]
    # the reference computes `angle * M_PI/180.0f` in double then rounds to float
    st = [float(np.float32(float(np.float32(a)) * np.pi / float(np.float32(180.0)))) for a in (-30, -15, 0, 15, 30)]
    return make_params(shot_interval=300, shot_decay=10, resolution=0.5, obstacle_threshold=0.75,
                       prob_min=0.1, prob_max=0.95, prob_free=0.4, grid_size=60, allow_diag=1,
                       step_size=0.75, max_lat_acc=4.0, max_long_dec=2.0, wheelbase=2.269,
                       rear_to_cg=1.1, apf_rep_constant=1.0,
                       apf_active_angle=float(np.float32(np.pi / 4)), num_angle_bins=72,
                       num_actions=1, steering=st, curvature_weights=[0.0] * 5)


REF_TEST_LINES = np.array([[21.9, 4.5, 21.9, 31.5], [20.4, 33.0, 38.4, 33.0],
                           [10.5, 4.5, 10.5, 40.5], [9.0, 42.0, 39.0, 42.0]], np.float32)
REF_TEST_BOXES = np.array([[18.0, 22.8, 3.5, 2.9], [14.25, 28.5, 2.0, 5.3], [18.0, 34.8, 3.5, 2.9]], np.float32)
REF_TEST_START = np.array([18.0, 18.0, np.pi / 2], np.float32)
REF_TEST_GOAL = np.array([26.0, 36.0, 0.0], np.float32)


def setup_ref_test_scenario(o):
    """Apply the map-building calls of utils/hybrid_astar/test_hybrid_astar.cpp:77-84 to planner `o`."""
    o.update_goal(REF_TEST_GOAL, REF_TEST_START)
    for _ in range(5):
        o.decay()
        o.update_lines(REF_TEST_LINES, np.full(4, 0.6, np.float32), 1.25)
        o.update_boxes(REF_TEST_BOXES, np.full(3, 0.75, np.float32), 2.5)


def field2d(o):
    """Double-precision Dijkstra distance-to-goal field of the port oracle (negative = unreachable)."""
    out = np.empty((o.N, o.N), np.float64)
    o._fn("field2d")(o.h, _fp(out))
    return out
