"""GPU test of the C++ drop-in API: tests/cpp/api_driver.cpp (built by path_planning_pkg_b200.build) drives
HybridAStar<float/double>, Dubins, VehicleModel, Grid3D, AStar, VelocityGenerator and PedestrianHandler through
include/path_planning_pkg/*.h + libpath_planning_b200.so and prints results that are checked against the reference's
recorded vectors and the compiled reference."""
import os
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
BIN = os.path.join(orc.ROOT, "tests", "cpp", "bin", "api_driver")


@pytest.fixture(scope="module")
def output():
    assert os.path.exists(BIN), "build tests/cpp/bin/api_driver first (path_planning_pkg_b200.build.build_cpp_tests)"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout.split("\n")


def _line(lines, prefix):
    return [l for l in lines if l.startswith(prefix)]


def test_hybrid_astar_float_matches_golden(output):
    head = _line(output, "f32 success")[0].split()
    assert head[2] == "1" and abs(float(head[4]) - 33.0305) < 1e-4
    assert head[6] == "43" and head[8] == "43" and int(head[10]) == 882
    pts = np.array([[float(v) for v in l.split()[2:]] for l in _line(output, "f32 pt")])
    gold = np.load(orc.ROOT + "/tests/golden/hybrid_astar_path.npy")
    assert pts.shape == gold.shape and np.allclose(pts, gold, rtol=2e-5, atol=2e-5)
    if orc.have_ref():
        o = orc.ref(orc.ref_test_params())
        orc.setup_ref_test_scenario(o)
        b = o.find_path(2.0, orc.REF_TEST_START)
        assert np.allclose(pts, b["path"][::-1], rtol=0, atol=5e-6)     # printed with 6 decimals
        thr = o.consts().log_threshold
        assert int(head[12]) == int((o.get_map() >= thr).sum())


def test_hybrid_astar_double_api(output):
    """T = double converts at the boundary and computes in FP32 on the device: same plan as the float run."""
    f32 = _line(output, "f32 success")[0].split(); f64 = _line(output, "f64 success")[0].split()
    assert f64[2] == "1" and abs(float(f64[4]) - float(f32[4])) < 1e-3 and f64[6] == f32[6]


def test_failure_contract(output):
    for tag in ("f32", "f64"):
        b = _line(output, tag + " blocked")[0].split()
        assert b[3] == "0" and b[5] == "1" and b[7] == "0"      # {max, false}, vectors untouched


def test_velocity_profile(output):
    v = _line(output, "f32 velocity")[0].split()
    assert v[3] == "1" and int(v[5]) == 43 and abs(float(v[9])) < 1e-6     # feasible, stops at the goal


def test_dubins_vehicle_grid_astar(output):
    d = _line(output, "dubins")[0].split()
    assert abs(float(d[2]) - 4.08106) < 1e-4 and abs(float(d[4]) - 36.830315) < 1e-4 and d[6] == "RSL" and d[8] == "73"
    v = _line(output, "vehicle")[0].split()
    assert v[2] == "3" and v[4] == "0" and v[6] == "3" and v[8] == "1"
    gold = np.load(orc.ROOT + "/tests/golden/vehicle_rollout.npy")
    assert abs(float(v[10]) - gold[1, 0]) < 4e-3 and abs(float(v[12]) - gold[1, 1]) < 4e-3
    g = _line(output, "grid3d")[0].split()
    assert int(g[2]) == 23 and int(g[7]) >= 1 and float(g[11]) > 10.0 and g[11] == g[13]     # cached value returned again
    p = _line(output, "pedestrian")[0].split()
    assert float(p[2]) >= 0.0
