"""Generic vehicle-footprint collision kernel (north_star (c); pp_set_footprint / pp_footprint_batch) on the GPU: tables,
free flags, cells and blocked-cell counts equal the statement of the semantics (oracle/port/footprint.inc) bit for bit, the
zero-size footprint equals the one-cell kernel that mirrors the reference (pp_collision_batch), edge cases included."""
import json
import os

import numpy as np
import pytest

import orc
import scenarios as S
from test_cpu_footprint import VEHICLES, poses
from test_gpu_parity import _ctx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    sc = S.c4_group(2, n_starts=4)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    ctx, o = _ctx(P), orc.port(P)
    for x in (ctx, o):
        S.build_map(x, sc)
    assert np.array_equal(ctx.get_map().view(np.uint32), o.get_map().view(np.uint32))
    return ctx, o, P


@pytest.mark.parametrize("veh", VEHICLES)
def test_kernel_equals_statement(pair, veh):
    ctx, o, P = pair
    ctx.set_footprint(*veh)
    for b in (0, 1, 9, 18, 36, 54, 71, 72):
        assert np.array_equal(ctx.footprint_table(b), o.footprint_table(b, *veh)), (veh, b)
    p = poses(P, 20000, 11)
    fa, ca, ha = ctx.footprint(p)
    fb, cb, hb = o.footprint_check(p, *veh)
    assert np.array_equal(fa, fb) and np.array_equal(ca, cb) and np.array_equal(ha, hb)
    assert 0 < fa.sum() < len(fa)


def test_zero_size_footprint_equals_the_one_cell_kernel(pair):
    ctx, o, P = pair
    ctx.set_footprint(0.0, 0.0, 0.0)
    p = poses(P, 20000, 5)
    free, cells, hits = ctx.footprint(p)
    f1, c1 = ctx.collision(p[:, :2].copy())
    assert np.array_equal(free, f1) and np.array_equal(cells, c1) and np.array_equal(hits, 1 - free)


def test_edge_cases(pair):
    ctx, o, P = pair
    ctx.set_footprint(4.0, 2.0, 1.0)
    L = P.grid_size * P.resolution
    p = np.array([[0.0, 0.0, 0.0], [L - 0.01, L - 0.01, 0.7], [-0.1, 5.0, 0.0], [L / 2, -3.0, 1.0], [L + 5, L / 2, -2.0],
                  [L / 2, L / 2, np.pi], [L / 2, L / 2, -np.pi]], np.float32)          # borders, outside, bin 72 / bin 0
    fa, ca, ha = ctx.footprint(p)
    fb, cb, hb = o.footprint_check(p, 4.0, 2.0, 1.0)
    assert np.array_equal(fa, fb) and np.array_equal(ca, cb) and np.array_equal(ha, hb)
    assert fa[0] == 0 and fa[4] == 0                                                   # the rectangle leaves the grid
    one = ctx.footprint(p[:1])                                                         # n = 1
    assert one[0][0] == fa[0]
    with pytest.raises(RuntimeError):
        ctx.set_footprint(40.0, 2.0, 1.0)                                              # 200 cells: beyond the supported window


def test_kernel_rate(pair):
    """Not a pass/fail number: records the kernel's rate for profiles/ when run on the GPU box."""
    ctx, o, P = pair
    ctx.set_footprint(4.0, 2.0, 1.0)
    p = poses(P, 1 << 20, 3)
    ctx.footprint(p)
    best = min(ctx.footprint(p, want_ms=True)[3] for _ in range(5))
    cells = float(np.mean([len(ctx.footprint_table(b)) for b in range(P.num_angle_bins)]))
    variant = "staged" if os.environ.get("PP_B200_FOOT_STAGED") == "1" else "direct"
    out = {"kernel": f"pp_footprint_kernel<{variant}>", "poses": len(p), "vehicle_m": [4.0, 2.0, 1.0], "grid": [P.grid_size, P.resolution],
           "mean_footprint_cells": cells, "kernel_ms": best, "poses_per_s": len(p) / (best * 1e-3),
           "algorithmic_GBps": len(p) * (cells * 4 + 12 + 12) / (best * 1e-3) / 1e9}
    d = os.path.join(orc.ROOT, "gpurun_out")
    if os.path.isdir(d):
        json.dump(out, open(os.path.join(d, f"footprint_kernel_{variant}.json"), "w"))
    assert best > 0
