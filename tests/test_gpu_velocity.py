"""SURVEY.md §8(f) N3 on the GPU: pp_velocity_profile_batch (one thread per path) and pp_trajectory_batch (one warp per query of
the last batch: reconstruct_path + velocity profile + /local_planner/trajectory layout on the device) against the unmodified
reference's VelocityGenerator / HybridAStar through the compiled oracle, bit for bit (NaN results NaN in the same places)."""
import numpy as np
import pytest

import orc
import scenarios as S
from test_cpu_velocity import FLT_MAX, LIM, random_paths, same_bits_nan_aware
from test_gpu_parity import _ctx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("signed", [False, True])
def test_profile_kernel_equals_reference(signed):
    n, cap = 300, 128
    xy, cv, cnt, vi, vc, fl = random_paths(n, cap, 21 + signed, signed)
    ctx = _ctx(orc.make_params(grid_size=16, resolution=0.5))
    vel, ok = ctx.velocity_profile_batch(LIM, xy, cv, cnt, vi, vc, fl)
    ref = orc.ref(orc.make_params(grid_size=16, resolution=0.5))
    for k in range(n):
        m = cnt[k]
        xyh = np.concatenate([xy[k, :m], np.zeros((m, 1), np.float32)], 1)
        rv, rok = ref.velocity_profile(LIM, float(vi[k]), float(vc[k]), xyh, cv[k, :m], coast=bool(fl[k] & 1), stop=bool(fl[k] & 2))
        assert same_bits_nan_aware(vel[k, :m], rv), k
        assert bool(ok[k]) == rok, k
    # optional arguments: no cap, no flags
    vel2, ok2 = ctx.velocity_profile_batch(LIM, xy[:8], cv[:8], cnt[:8], vi[:8])
    for k in range(8):
        m = cnt[k]
        xyh = np.concatenate([xy[k, :m], np.zeros((m, 1), np.float32)], 1)
        rv, rok = ref.velocity_profile(LIM, float(vi[k]), FLT_MAX, xyh, cv[k, :m])
        assert same_bits_nan_aware(vel2[k, :m], rv) and bool(ok2[k]) == rok


@pytest.mark.parametrize("mode", [0, 1])
def test_trajectory_batch_equals_reference_pipeline(mode):
    """EXACT mode: path and profile equal the reference's classes end to end.  K-POP mode (own path semantics): the device
    trajectory equals the reference's VelocityGenerator applied to the path the C ABI returns for the same query."""
    sc = S.c1_scenario(2)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, o = _ctx(P), orc.ref(P)
    for x in (ctx, o):
        S.build_map(x, sc)
    starts = np.array([[0.0, 0.0, 0.0, 3.0], [1.0, 0.5, 0.1, 1.0], [0.5, -1.0, -0.2, 0.0], [2.0, 1.0, 0.3, 2.0],
                       [-400.0, 0.0, 0.0, 1.0]], np.float32)                     # the last one starts outside the grid
    q = ctx.make_queries(starts, [0] * len(starts))
    res, paths, curv, _ = ctx.find_path_batch(q, ctx.make_opts(path_cap=1024, mode=mode, kpop=32))
    vcap = np.array([FLT_MAX, 2.5, 4.0, FLT_MAX, 3.0], np.float32); stop = np.array([0, 1, 1, 0, 1], np.int32)
    trajs, ok = ctx.trajectory_batch(LIM, vcap, stop)
    checked = 0
    for k in range(len(starts)):
        if not res[k]["success"]:
            assert trajs[k].shape == (4, 0) and ok[k] == 0
            continue
        m = int(res[k]["n_path"])
        path, cv = paths[k, :m], curv[k, :m]
        if mode == 0:
            o.scrub()
            b = o.find_path(float(starts[k, 3]), starts[k, :3])
            if b["n_pops_bin_oob"] == 0:
                assert b["success"] and np.array_equal(path.view(np.uint32), b["path"].view(np.uint32))
        rv, rok = o.velocity_profile(LIM, float(starts[k, 3]), float(vcap[k]), path, cv, coast=False, stop=bool(stop[k]))
        want = np.stack([path[::-1, 0], path[::-1, 1], path[::-1, 2], rv])
        assert trajs[k].shape == (4, m) and same_bits_nan_aware(trajs[k], want), k
        assert bool(ok[k]) == rok
        checked += 1
    assert checked >= 3
    # optional arguments absent: no cap, never stop
    trajs2, ok2 = ctx.trajectory_batch(LIM)
    k = 0
    m = int(res[k]["n_path"])
    rv, rok = o.velocity_profile(LIM, float(starts[k, 3]), FLT_MAX, paths[k, :m], curv[k, :m])
    assert same_bits_nan_aware(trajs2[k][3], rv) and bool(ok2[k]) == rok
