"""SURVEY.md §8(f) N3 on the CPU: the device code of the velocity profile and the trajectory assembly
(path_planning_pkg_b200/csrc/core/pp_velocity.h, here in the host-lane build tests/cpp/host_emul.cpp) against the unmodified
reference: VelocityGenerator<float>::generate_velocity_profile (lib/VelocityGenerator.cpp:19-85) on random paths, and
find_path + generate_velocity_profile + the message layout of LocalPlanner::publish_trajectory (src/local_planner.cpp:346-372)
end to end.  Bit for bit; NaN results (the reference roots a negative number when v^2 * curvature exceeds the lateral limit) must
be NaN in the same places."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
import scenarios as S

EMU_SO = os.path.join(orc.ROOT, "tests", "cpp", "bin", "libpp_host_emul.so")
LIM = [5.0, 1.0, 2.0, 1.0, 2.5]                    # launch/local_planner.launch:25-29
FLT_MAX = float(np.finfo(np.float32).max)

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="needs the compiled reference")


def same_bits_nan_aware(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def random_paths(n, cap, seed, signed=False):
    """paths in find_path's order (goal -> start): piecewise-constant curvature arcs, |curvature| like the caller's."""
    rs = np.random.RandomState(seed)
    xy = np.zeros((n, cap, 2), np.float32); cv = np.zeros((n, cap), np.float32); cnt = np.zeros(n, np.int32)
    for k in range(n):
        m = 1 + k if k < 4 else rs.randint(2, cap + 1)
        ds = rs.uniform(0.2, 0.6); x, y, h, kap = rs.uniform(-5, 5), rs.uniform(-5, 5), rs.uniform(-3, 3), 0.0
        pts, cs = [], []
        for i in range(m):
            if i % 7 == 0:
                kap = 0.0 if k % 5 == 0 else rs.uniform(-0.24, 0.24)
            if i % 11 == 5:
                kap = 0.0
            pts.append((x, y)); cs.append(kap if signed else abs(kap))
            x += ds * np.cos(h); y += ds * np.sin(h); h += ds * kap
        xy[k, :m] = np.array(pts[::-1], np.float32); cv[k, :m] = np.array(cs[::-1], np.float32); cnt[k] = m
    vi = rs.uniform(0, 2.4, n).astype(np.float32)
    vc = np.where(np.arange(n) % 3 == 0, FLT_MAX, rs.uniform(0.5, 6, n)).astype(np.float32)
    fl = (np.arange(n) % 4).astype(np.int32)
    return xy, cv, cnt, vi, vc, fl


@pytest.fixture(scope="module")
def emu_lib(built):
    return C.CDLL(EMU_SO)


@pytest.mark.parametrize("signed", [False, True])
def test_profile_core_equals_reference(emu_lib, signed):
    n, cap = 300, 128
    xy, cv, cnt, vi, vc, fl = random_paths(n, cap, 21 + signed, signed)
    vel = np.zeros((n, cap), np.float32); ok = np.zeros(n, np.int32); lim = np.asarray(LIM, np.float32)
    emu_lib.emu_velocity_profile_batch(orc._fp(lim), orc._fp(xy), orc._fp(cv), orc._fp(cnt), C.c_int(n), C.c_int(cap), orc._fp(vi),
                                       orc._fp(vc), orc._fp(fl), orc._fp(vel), orc._fp(ok))
    ref = orc.ref(orc.make_params(grid_size=16, resolution=0.5))
    nan_rows = 0
    for k in range(n):
        m = cnt[k]
        xyh = np.concatenate([xy[k, :m], np.zeros((m, 1), np.float32)], 1)
        rv, rok = ref.velocity_profile(LIM, float(vi[k]), float(vc[k]), xyh, cv[k, :m], coast=bool(fl[k] & 1), stop=bool(fl[k] & 2))
        assert same_bits_nan_aware(vel[k, :m], rv), k
        assert bool(ok[k]) == rok, k
        nan_rows += int(np.isnan(rv).any())
    assert (nan_rows > 10) if signed else (nan_rows < n // 4)


@pytest.mark.parametrize("seed", [0, 2, 5])
def test_trajectory_equals_reference_pipeline(emu_lib, seed):
    """find_path -> velocity profile -> /local_planner/trajectory layout, device code vs the reference's classes."""
    sc = S.c1_scenario(seed)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    e, o = orc.Oracle(emu_lib, "emu", P), orc.ref(P)
    for x in (e, o):
        S.build_map(x, sc)
    for vel, vcap, stop in [(3.0, FLT_MAX, False), (1.0, 2.5, True), (0.0, 4.0, True)]:
        q = sc["queries"][0]
        o.scrub()
        a, b = e.find_path(vel, q[:3]), o.find_path(vel, q[:3])
        assert b["success"] and a["n_pops"] == b["n_pops"]
        traj, ok = e.trajectory(LIM, vcap, stop)
        rv, rok = o.velocity_profile(LIM, vel, vcap, b["path"], b["curvature"], coast=False, stop=stop)
        m = len(b["path"])
        want = np.stack([b["path"][::-1, 0], b["path"][::-1, 1], b["path"][::-1, 2], rv])      # publish_trajectory, :364-368
        assert traj.shape == (4, m) and same_bits_nan_aware(traj, want)
        assert ok == rok
        if stop:
            assert traj[3, -1] == 0.0
