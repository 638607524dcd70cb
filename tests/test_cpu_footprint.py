"""Generic vehicle-footprint collision check (north_star (c)) on the CPU: the product's table builder and
per-cell predicates (host-lane build, tests/cpp/host_emul.cpp) against the statement of the semantics in
oracle/port/footprint.inc, and the zero-size footprint against the reference's own collision check."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
import scenarios as S

EMU_SO = os.path.join(orc.ROOT, "tests", "cpp", "bin", "libpp_host_emul.so")
VEHICLES = [(0.0, 0.0, 0.0), (4.0, 2.0, 1.0), (2.5, 1.2, 0.4), (0.3, 0.0, 0.0), (6.0, 2.6, 3.0)]


@pytest.fixture(scope="module")
def pair(built):
    sc = S.c4_group(2, n_starts=4)
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    e, o = orc.Oracle(C.CDLL(EMU_SO), "emu", P), orc.port(P)
    for x in (e, o):
        S.build_map(x, sc)
    return e, o, P


def poses(P, n, seed):
    rs = np.random.RandomState(seed)
    L = P.grid_size * P.resolution
    xy = rs.uniform(-0.05 * L, 1.05 * L, (n, 2))                   # some poses outside the grid
    xy[: n // 2] = rs.uniform(0.3 * L, 0.75 * L, (n // 2, 2))      # half of them inside the clutter corridor
    h = rs.uniform(-np.pi, np.pi, (n, 1))
    h[:8, 0] = [np.pi, -np.pi, 3.12, -3.12, 0.0, np.pi / 2, -np.pi / 2, 3.1415]       # bin 72 / bin 0 edge
    return np.concatenate([xy, h], 1).astype(np.float32)


@pytest.mark.parametrize("veh", VEHICLES)
def test_table_equals_statement(pair, veh):
    e, o, P = pair
    total = 0
    for b in range(P.num_angle_bins + 1):
        a, r = e.footprint_table(b, *veh), o.footprint_table(b, *veh)
        assert np.array_equal(a, r), (veh, b)
        total += len(a)
    if veh == (0.0, 0.0, 0.0):
        assert total == P.num_angle_bins + 1 and not e.footprint_table(0, *veh).any()       # {(0, 0)} in every bin
    else:
        assert total > 2 * (P.num_angle_bins + 1)


@pytest.mark.parametrize("veh", VEHICLES)
def test_check_equals_statement(pair, veh):
    e, o, P = pair
    p = poses(P, 4000, 11)
    fa, ca, ha = e.footprint_check(p, *veh)
    fb, cb, hb = o.footprint_check(p, *veh)
    assert np.array_equal(fa, fb) and np.array_equal(ca, cb) and np.array_equal(ha, hb)
    assert 0 < fa.sum() < len(fa)


def test_zero_size_footprint_is_the_references_check(pair):
    """{(0, 0)} = Grid3D::get_neighbors' filter (Grid3D.cpp:53-59): same booleans and cells as the reference keeps."""
    e, o, P = pair
    p = poses(P, 3000, 5)
    free, cells, hits = e.footprint_check(p, 0.0, 0.0, 0.0)
    m = o.get_map(); thr = o.consts().log_threshold; N = P.grid_size
    ci = (p[:, 0] / np.float32(P.resolution)).astype(np.int32); cj = (p[:, 1] / np.float32(P.resolution)).astype(np.int32)
    inside = (ci > -1) & (ci < N) & (cj > -1) & (cj < N)
    want = np.zeros(len(p), bool)
    want[inside] = m[ci[inside], cj[inside]] < thr
    assert np.array_equal(cells[:, 0], ci) and np.array_equal(cells[:, 1], cj)
    assert np.array_equal(free.astype(bool), want) and np.array_equal(hits, 1 - free)


def test_footprint_is_monotone_in_vehicle_size(pair):
    """A larger rectangle can only block more poses; every blocked one-cell pose stays blocked."""
    e, o, P = pair
    p = poses(P, 3000, 7)
    f0 = e.footprint_check(p, 0.0, 0.0, 0.0)[0]
    f1 = e.footprint_check(p, 2.5, 1.2, 0.4)[0]
    f2 = e.footprint_check(p, 4.0, 2.0, 1.0)[0]
    assert np.all(f1 <= f0) and f2.sum() < f1.sum() < f0.sum()
