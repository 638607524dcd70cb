"""ctypes binding of lib/libpp_b200.so (C ABI in include/pp_b200.h).

This is harness plumbing for tests and bench.py: the product is the CUDA library and the C++ classes
above it.  The binding fails loudly when the library is missing or when there is no CUDA device --
there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpp_b200.so")
MAX_STEER = 16


class PPError(RuntimeError):
    pass


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


class ConstsInfo(C.Structure):
    _fields_ = [("log_threshold", C.c_float), ("log_min", C.c_float), ("log_max", C.c_float), ("log_free", C.c_float),
                ("precision", C.c_float), ("r_min", C.c_float), ("ang_step", C.c_float), ("n2", C.c_int), ("n45", C.c_int)]


class FrameInfo(C.Structure):
    _fields_ = [("grid_heading", C.c_float), ("goal_world", C.c_float * 3), ("goal_grid", C.c_float * 3),
                ("goal_bin", C.c_int), ("goal_ci", C.c_int), ("goal_cj", C.c_int), ("num_apf", C.c_int)]


class SearchOpts(C.Structure):
    _fields_ = [("max_expansions", C.c_int), ("max_open", C.c_int), ("max_open2d", C.c_int),
                ("path_cap", C.c_int), ("trace_cap", C.c_int), ("max_slots", C.c_int),
                ("mode", C.c_int), ("kpop", C.c_int)]


STATE_DT = np.dtype([("x", "f4"), ("y", "f4"), ("heading", "f4"), ("g", "f4"), ("f", "f4"),
                     ("vmin_sqr", "f4"), ("curvature_index", "i4"), ("angle_bin", "i4"),
                     ("ci", "i4"), ("cj", "i4")])
POP_DT = np.dtype([("ci", "i4"), ("cj", "i4"), ("bin", "i4"), ("x", "f4"), ("y", "f4"),
                   ("heading", "f4"), ("g", "f4"), ("f", "f4")])
QUERY_DT = np.dtype([("x", "f4"), ("y", "f4"), ("heading", "f4"), ("vel", "f4"), ("group", "i4")])
RESULT_DT = np.dtype([("success", "i4"), ("status", "i4"), ("cost", "f4"), ("n_pops", "i4"),
                      ("n_pops_bin_oob", "i4"), ("n_chain", "i4"), ("n_dubins", "i4"),
                      ("n_lazy_searches", "i4"), ("n_lazy_pops", "i4"), ("max_open", "i4"),
                      ("n_closed", "i4"), ("n_path", "i4")])

_lib = None


def load():
    """Load libpp_b200.so or raise: the product path must not degrade silently."""
    global _lib
    if _lib is None:
        path = os.environ.get("PP_B200_LIB", LIB_PATH)   # development override (e.g. the -DPP_PROFILE variant)
        if not os.path.exists(path):
            raise PPError(f"{path} is missing: build it with `python -m path_planning_pkg_b200.build` "
                          "(there is no CPU fallback)")
        lib = C.CDLL(path)
        lib.pp_last_error.restype = C.c_char_p
        lib.pp_map_device_ptr.restype = C.c_void_p
        lib.pp_kernel_launches.restype = C.c_ulonglong
        _lib = lib
    return _lib


def exported_symbols():
    """Names declared in include/pp_b200.h (used by the CPU test that checks the library exports them all)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "pp_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", txt)))


def make_params(**kw):
    """Launch defaults of the reference (launch/local_planner.launch:11-45), overridable by keyword."""
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
        if k not in ("steering", "curvature_weights"):
            setattr(p, k, v)
    st = d["steering"]
    cw = list(d["curvature_weights"]) + [0.0] * MAX_STEER
    p.num_steering = len(st)
    for i, a in enumerate(st):
        p.steering[i] = a
        p.curvature_weights[i] = cw[i]
    return p


def params_from(other):
    """Copy any ctypes struct with the same field layout (e.g. the oracle's) into Params."""
    p = Params()
    C.memmove(C.byref(p), C.byref(other), C.sizeof(Params))
    return p


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """pp_context: shared parameters + `num_groups` planner instances (map, goal frame, APF list)."""

    def __init__(self, params, num_groups=1, device=0, _lane_of=None):
        self.lib = load()
        self.params = params
        self.num_groups = num_groups
        self.N = params.grid_size
        self.stride = 2 * params.num_actions + 1
        self._parent = _lane_of          # keeps the parent alive as long as its lane
        h = C.c_void_p()
        if _lane_of is None:
            self._chk(self.lib.pp_create(C.byref(params), C.c_int(device), C.c_int(num_groups), C.byref(h)))
        else:
            self._chk(self.lib.pp_create_lane(_lane_of.h, C.byref(h)))
        self.h = h

    def create_lane(self):
        """pp_create_lane: a context sharing this one's maps / frames / APF lists with its own stream and search scratch."""
        return Context(self.params, self.num_groups, _lane_of=self)

    def set_memory_budget(self, nbytes):
        self._chk(self.lib.pp_set_memory_budget(self.h, C.c_ulonglong(int(nbytes))))

    def _chk(self, rc):
        if rc != 0:
            raise PPError(f"pp_b200 error {rc}: {self.lib.pp_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map / frame ----
    def update_goal(self, goal, start, group=0):
        g = np.asarray(goal, np.float32); s = np.asarray(start, np.float32)
        self._chk(self.lib.pp_update_goal(self.h, C.c_int(group), _p(g), _p(s)))

    def reset(self, group=0):
        self._chk(self.lib.pp_reset(self.h, C.c_int(group)))

    def set_history(self, group=0, enable=True):
        """Planner-object semantics for `group` (pp_set_history): single EXACT-mode queries continue on the 2D heuristic
        cache earlier single queries left, like successive find_path calls on one reference HybridAStar object."""
        self._chk(self.lib.pp_set_history(self.h, C.c_int(group), C.c_int(1 if enable else 0)))

    def update_boxes(self, boxes, conf, apf_added_radius, group=0):
        b = np.ascontiguousarray(boxes, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._chk(self.lib.pp_update_obstacles_boxes(self.h, C.c_int(group), _p(b), _p(c), C.c_int(len(c)),
                                                     C.c_float(apf_added_radius)))

    def update_boxes_2d(self, boxes, conf, group=0):
        b = np.ascontiguousarray(boxes, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._chk(self.lib.pp_update_obstacles_boxes_2d(self.h, C.c_int(group), _p(b), _p(c), C.c_int(len(c))))

    def update_boxes_2d_decay(self, boxes, conf, group=0):
        """pp_update_obstacles_boxes_2d_decay: boxes + whole-map decay in one pass (asynchronous)."""
        b = np.ascontiguousarray(boxes, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._chk(self.lib.pp_update_obstacles_boxes_2d_decay(self.h, C.c_int(group), _p(b), _p(c), C.c_int(len(c))))

    def update_lines(self, lines, conf, width, group=0):
        b = np.ascontiguousarray(lines, np.float32); c = np.ascontiguousarray(conf, np.float32)
        self._chk(self.lib.pp_update_obstacles_lines(self.h, C.c_int(group), _p(b), _p(c), C.c_int(len(c)), C.c_float(width)))

    def decay(self, group=0):
        self._chk(self.lib.pp_update_obstacles_decay(self.h, C.c_int(group)))

    def get_map(self, group=0):
        out = np.empty((self.N, self.N), np.float32)
        self._chk(self.lib.pp_map_download(self.h, C.c_int(group), _p(out)))
        return out

    def set_map(self, m, group=0):
        m = np.ascontiguousarray(m, np.float32)
        assert m.shape == (self.N, self.N)
        self._chk(self.lib.pp_map_upload(self.h, C.c_int(group), _p(m)))

    def map_device_ptr(self, group=0):
        return self.lib.pp_map_device_ptr(self.h, C.c_int(group))

    def map_mark_dirty(self, group=0):
        self._chk(self.lib.pp_map_mark_dirty(self.h, C.c_int(group)))

    def update_apf(self, boxes, apf_added_radius, group=0):
        """pp_update_obstacles_apf: the APF obstacle list only (the map itself arrives by broadcast_maps)."""
        b = np.ascontiguousarray(boxes, np.float32)
        self._chk(self.lib.pp_update_obstacles_apf(self.h, C.c_int(group), _p(b), C.c_int(len(b)), C.c_float(apf_added_radius)))

    # ---- map replication over NCCL (one communicator per context, one rank per GPU) ----
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self._chk(self.lib.pp_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._chk(self.lib.pp_comm_init(self.h, C.c_int(nranks), C.c_int(rank), buf))

    def broadcast_maps(self, first_group=0, n_groups=None, root=0):
        n = self.num_groups - first_group if n_groups is None else n_groups
        self._chk(self.lib.pp_broadcast_maps(self.h, C.c_int(first_group), C.c_int(n), C.c_int(root)))

    def sync(self):
        self._chk(self.lib.pp_sync(self.h))

    def consts(self):
        c = ConstsInfo()
        self._chk(self.lib.pp_get_consts(self.h, C.byref(c)))
        return c

    def frame(self, group=0):
        f = FrameInfo()
        self._chk(self.lib.pp_get_frame(self.h, C.c_int(group), C.byref(f)))
        return f

    def tables(self):
        S, B = self.params.num_steering, self.params.num_angle_bins
        oxy = np.empty((S, B, 2), np.float32); oh = np.empty(S, np.float32)
        ac = np.empty(S, np.float32); cu = np.empty(S, np.float32)
        self._chk(self.lib.pp_get_tables(self.h, _p(oxy), _p(oh), _p(ac), _p(cu)))
        return oxy, oh, ac, cu

    def kernel_launches(self):
        return int(self.lib.pp_kernel_launches(self.h))

    def batch_retried(self):
        return int(self.lib.pp_batch_retried(self.h))

    # ---- stateless batches ----
    def set_start(self, queries):
        q = np.ascontiguousarray(queries, QUERY_DT)
        out = np.zeros(len(q), STATE_DT)
        self._chk(self.lib.pp_set_start_batch(self.h, _p(q), C.c_int(len(q)), _p(out)))
        return out

    def _succ(self, fn, states, group=None):
        states = np.ascontiguousarray(states, STATE_DT)
        n = len(states)
        out = np.zeros((n, self.stride), STATE_DT)
        cnt = np.zeros(n, np.int32); fl = np.zeros(n, np.int32)
        if group is None:
            self._chk(fn(self.h, _p(states), C.c_int(n), _p(out), _p(cnt), _p(fl)))
        else:
            self._chk(fn(self.h, C.c_int(group), _p(states), C.c_int(n), _p(out), _p(cnt), _p(fl)))
        return out, cnt, fl

    def rollout(self, states):
        return self._succ(self.lib.pp_rollout_batch, states)

    def expand(self, states, group=0):
        return self._succ(self.lib.pp_expand_batch, states, group)

    def collision(self, xy, group=0):
        xy = np.ascontiguousarray(xy, np.float32)
        n = len(xy)
        fr = np.zeros(n, np.int32); cells = np.zeros((n, 2), np.int32)
        self._chk(self.lib.pp_collision_batch(self.h, C.c_int(group), _p(xy), C.c_int(n), _p(fr), _p(cells)))
        return fr.astype(bool), cells

    def set_footprint(self, length, width, rear_overhang):
        """Per-heading-bin cell offsets of the vehicle rectangle (pp_set_footprint); (0, 0, 0) = the reference's one-cell check."""
        self._chk(self.lib.pp_set_footprint(self.h, C.c_float(length), C.c_float(width), C.c_float(rear_overhang)))

    def footprint_table(self, bin_, cap=4096):
        n = C.c_int(); offs = np.zeros((cap, 2), np.int16)
        self._chk(self.lib.pp_get_footprint(self.h, C.c_int(bin_), C.byref(n), _p(offs), C.c_int(cap)))
        return offs[:n.value].copy()

    def footprint(self, xyh, group=0, want_ms=False):
        """pp_footprint_batch: free flags, base cells, blocked-cell counts for n grid-frame poses (x, y, heading)."""
        p = np.ascontiguousarray(xyh, np.float32); n = len(p)
        free = np.zeros(n, np.int32); cells = np.zeros((n, 2), np.int32); hits = np.zeros(n, np.int32); ms = C.c_float()
        self._chk(self.lib.pp_footprint_batch(self.h, C.c_int(group), _p(p), C.c_int(n), _p(free), _p(cells), _p(hits), C.byref(ms)))
        return (free, cells, hits, ms.value) if want_ms else (free, cells, hits)

    def check_path(self, xyh, group=0):
        xyh = np.ascontiguousarray(xyh, np.float32)
        fr = C.c_int()
        self._chk(self.lib.pp_check_path(self.h, C.c_int(group), _p(xyh), C.c_int(len(xyh)), C.byref(fr)))
        return bool(fr.value)

    def apf(self, xyh, group=0):
        xyh = np.ascontiguousarray(xyh, np.float32)
        out = np.empty(len(xyh), np.float32)
        self._chk(self.lib.pp_apf_batch(self.h, C.c_int(group), _p(xyh), C.c_int(len(xyh)), _p(out)))
        return out

    def dubins_length(self, starts, goal):
        starts = np.ascontiguousarray(starts, np.float32); goal = np.asarray(goal, np.float32)
        n = len(starts)
        ln = np.empty(n, np.float32); ty = np.empty(n, np.int32); pr = np.empty((n, 4), np.float32)
        self._chk(self.lib.pp_dubins_length_batch(self.h, _p(starts), C.c_int(n), _p(goal), _p(ln), _p(ty), _p(pr)))
        return ln, ty, pr

    def dubins_length_fp32(self, starts, goal):
        starts = np.ascontiguousarray(starts, np.float32); goal = np.asarray(goal, np.float32)
        ln = np.empty(len(starts), np.float32)
        self._chk(self.lib.pp_dubins_length_fp32_batch(self.h, _p(starts), C.c_int(len(starts)), _p(goal), _p(ln)))
        return ln

    def dubins_path(self, start, goal, cap=4096):
        s = np.asarray(start, np.float32); g = np.asarray(goal, np.float32)
        xyh = np.empty((cap, 3), np.float32); cv = np.empty(cap, np.float32)
        n = C.c_int(); ln = C.c_float(); fl = C.c_int()
        self._chk(self.lib.pp_dubins_path(self.h, _p(s), _p(g), _p(xyh), _p(cv), C.c_int(cap), C.byref(n), C.byref(ln), C.byref(fl)))
        m = min(n.value, cap)
        return xyh[:m].copy(), cv[:m].copy(), ln.value, bool(fl.value)

    def astar_lazy(self, ij, group=0):
        ij = np.ascontiguousarray(ij, np.int32)
        out = np.empty(len(ij), np.float32)
        self._chk(self.lib.pp_astar_lazy_batch(self.h, C.c_int(group), _p(ij), C.c_int(len(ij)), _p(out)))
        return out

    # ---- batched search ----
    @staticmethod
    def make_queries(xyhv, groups=None):
        xyhv = np.asarray(xyhv, np.float32).reshape(-1, 4)
        q = np.zeros(len(xyhv), QUERY_DT)
        q["x"], q["y"], q["heading"], q["vel"] = xyhv[:, 0], xyhv[:, 1], xyhv[:, 2], xyhv[:, 3]
        q["group"] = 0 if groups is None else np.asarray(groups, np.int32)
        return q

    @staticmethod
    def make_opts(max_expansions=0, max_open=0, max_open2d=0, path_cap=0, trace_cap=0, max_slots=0, mode=0, kpop=0):
        """mode 0 = EXACT (reference-identical), 1 = KPOP (k pops per iteration, exact 2D field heuristic)."""
        return SearchOpts(max_expansions, max_open, max_open2d, path_cap, trace_cap, max_slots, mode, kpop)

    def find_path_batch(self, queries, opts=None, want_paths=True):
        """pp_find_path_batch with host buffers: H2D of the queries and D2H of results/paths inside the call."""
        q = np.ascontiguousarray(queries, QUERY_DT)
        n = len(q)
        opts = opts or self.make_opts()
        pc = opts.path_cap if opts.path_cap > 0 else 2048
        res = np.zeros(n, RESULT_DT)
        paths = np.zeros((n, pc, 3), np.float32) if want_paths else None
        curv = np.zeros((n, pc), np.float32) if want_paths else None
        trace = np.zeros((n, opts.trace_cap), POP_DT) if opts.trace_cap > 0 else None
        self._opts, self._n = opts, n
        self._chk(self.lib.pp_find_path_batch(self.h, _p(q), C.c_int(n), C.byref(opts), _p(res),
                                              _p(paths) if want_paths else None, _p(curv) if want_paths else None,
                                              _p(trace) if trace is not None else None))
        return res, paths, curv, trace

    def batch_upload(self, queries, opts=None):
        q = np.ascontiguousarray(queries, QUERY_DT)
        opts = opts or self.make_opts()
        self._opts, self._n = opts, len(q)
        self._chk(self.lib.pp_batch_upload(self.h, _p(q), C.c_int(len(q)), C.byref(opts)))

    def batch_run(self):
        ms = C.c_float()
        self._chk(self.lib.pp_batch_run(self.h, C.byref(ms)))
        return ms.value

    def batch_run_async(self):
        self._chk(self.lib.pp_batch_run_async(self.h))

    def batch_wait(self):
        ms = C.c_float()
        self._chk(self.lib.pp_batch_wait(self.h, C.byref(ms)))
        return ms.value

    def batch_fetch(self, want_paths=False):
        n, opts = self._n, self._opts
        pc = opts.path_cap if opts.path_cap > 0 else 2048
        res = np.zeros(n, RESULT_DT)
        paths = np.zeros((n, pc, 3), np.float32) if want_paths else None
        curv = np.zeros((n, pc), np.float32) if want_paths else None
        self._chk(self.lib.pp_batch_fetch(self.h, _p(res), _p(paths) if want_paths else None,
                                          _p(curv) if want_paths else None, None))
        return res, paths, curv

    def find_path(self, vel, start, group=0, pop_cap=1 << 17, path_cap=4096, **kw):
        """Single query with a pop trace: HybridAStar::find_path(vel_init, start, path, curvature)."""
        s = np.asarray(start, np.float32)
        q = self.make_queries([[s[0], s[1], s[2], vel]], [group])
        opts = self.make_opts(path_cap=path_cap, trace_cap=pop_cap, **kw)
        res, paths, curv, trace = self.find_path_batch(q, opts)
        r = res[0]
        n = int(r["n_path"])
        return dict(success=bool(r["success"]), cost=np.float32(r["cost"]), path=paths[0, :n].copy(),
                    curvature=curv[0, :n].copy(), pops=trace[0, :min(int(r["n_pops"]), pop_cap)].copy(),
                    n_pops=int(r["n_pops"]), n_pops_bin_oob=int(r["n_pops_bin_oob"]), status=int(r["status"]), raw=r)

    # ---- velocity profile / trajectory ----
    @staticmethod
    def _limits(lim5):
        return (C.c_float * 5)(*[float(v) for v in lim5])

    def velocity_profile_batch(self, lim5, paths_xy, curvature, counts, vel_init, vcap=None, flags=None):
        """pp_velocity_profile_batch: paths_xy [n][cap][2], curvature [n][cap] (goal -> start) -> velocity [n][cap], feasible [n]."""
        xy = np.ascontiguousarray(paths_xy, np.float32); cv = np.ascontiguousarray(curvature, np.float32)
        n, cap = cv.shape
        cnt = np.ascontiguousarray(counts, np.int32); vi = np.ascontiguousarray(vel_init, np.float32)
        vc = None if vcap is None else np.ascontiguousarray(vcap, np.float32)
        fl = None if flags is None else np.ascontiguousarray(flags, np.int32)
        vel = np.zeros((n, cap), np.float32); ok = np.zeros(n, np.int32)
        self._chk(self.lib.pp_velocity_profile_batch(self.h, self._limits(lim5), _p(xy), _p(cv), _p(cnt), C.c_int(n), C.c_int(cap), _p(vi),
                                                     _p(vc) if vc is not None else None, _p(fl) if fl is not None else None, _p(vel), _p(ok)))
        return vel, ok

    def trajectory_batch(self, lim5, vcap=None, stop=None, want_ms=False):
        """pp_trajectory_batch on the last batch: list of per-query (4, m) arrays [x, y, heading, velocity] start -> goal, feasible flags."""
        n, opts = self._n, self._opts
        pc = opts.path_cap if opts.path_cap > 0 else 2048
        traj = np.zeros((n, 4 * pc), np.float32); m = np.zeros(n, np.int32); ok = np.zeros(n, np.int32); ms = C.c_float()
        vc = None if vcap is None else np.ascontiguousarray(vcap, np.float32)
        st = None if stop is None else np.ascontiguousarray(stop, np.int32)
        self._chk(self.lib.pp_trajectory_batch(self.h, self._limits(lim5), _p(vc) if vc is not None else None,
                                               _p(st) if st is not None else None, _p(traj), _p(m), _p(ok), C.byref(ms)))
        out = [traj[k, :4 * m[k]].reshape(4, m[k]).copy() for k in range(n)]
        return (out, ok, ms.value) if want_ms else (out, ok)

    # ---- heuristic fields ----
    def field2d(self, group=0, download=True):
        """pp_heuristic_field_2d: exact 2D distance-to-goal field (FLT_MAX = unreachable); returns (field|None, sweeps, ms)."""
        out = np.empty((self.N, self.N), np.float32) if download else None
        sw = C.c_int(); ms = C.c_float()
        self._chk(self.lib.pp_heuristic_field_2d(self.h, C.c_int(group), _p(out) if download else None, C.byref(sw), C.byref(ms)))
        return out, sw.value, ms.value

    def field3d(self, group=0, use_h2d=True, download=True):
        """pp_heuristic_field_3d: max(h2d, Dubins) for every (i, j, heading bin); returns (field|None, ms)."""
        B = self.params.num_angle_bins
        out = np.empty((self.N, self.N, B), np.float32) if download else None
        ms = C.c_float()
        self._chk(self.lib.pp_heuristic_field_3d(self.h, C.c_int(group), C.c_int(1 if use_h2d else 0),
                                                 _p(out) if download else None, C.byref(ms)))
        return out, ms.value
