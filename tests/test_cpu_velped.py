"""SURVEY.md §8(f) rows N3 / N4 on the CPU: this repo's VelocityGenerator<T>::generate_velocity_profile and
PedestrianHandler<T>::calc_max_velocity (host classes in libpath_planning_b200.so, no GPU involved) against the
reference's (lib/VelocityGenerator.cpp:19-85, lib/PedestrianHandler.cpp:17-56), T = float and double, bit for bit:
512 velocity profiles (1..121 path points, coast / stop flags, with and without a pedestrian cap, straight paths, and
signed-curvature inputs whose NaN results must agree too) and 800 pedestrian scenes.  One driver source
(tests/cpp/velped_driver.cpp) is built against either library; both outputs must equal the committed digest
(tests/golden/velped_ref.json, made by tests/golden/make_velped_golden.py from the unmodified reference)."""
import json
import os
import subprocess
import sys

import pytest

import orc

sys.path.insert(0, os.path.join(orc.ROOT, "tests", "golden"))
from make_velped_golden import digest  # noqa: E402

GOLD = json.load(open(os.path.join(orc.ROOT, "tests", "golden", "velped_ref.json")))
REF_EXE = os.path.join(orc.ROOT, "oracle", "_ref", "velped_ref")


@pytest.fixture(scope="module")
def product_out():
    from path_planning_pkg_b200 import build
    build.build_cuda(verbose=False)
    build.build_host(verbose=False)
    exe = [b for b in build.build_cpp_tests(verbose=False) if b.endswith("velped_b200")][0]
    return subprocess.run([exe], capture_output=True, text=True, check=True).stdout


def test_product_matches_reference_digest(product_out):
    d = digest(product_out)
    assert d["lines"] == GOLD["lines"]
    bad = [k for k in GOLD["sections"] if d["sections"].get(k) != GOLD["sections"][k]]
    assert not bad, f"sections differing from the reference: {bad}; first lines {d['head'][:3]} vs {GOLD['head'][:3]}"
    assert d["sha256"] == GOLD["sha256"]


def test_cases_are_not_degenerate(product_out):
    lines = product_out.strip().split("\n")
    vel = [l for l in lines if l.startswith("f32 vel")]
    ped = [l.split(" : ")[1].strip() for l in lines if l.startswith("f32 ped")]
    finite = [l for l in vel if "ffc00000" not in l and "7fc00000" not in l]
    assert len(vel) == 256 and len(finite) > 200
    assert any(" ok 0 " in l for l in finite) and any(" ok 1 " in l for l in finite)
    # every branch of calc_max_velocity: no cap (FLT_MAX), stop (0), and a computed cap
    assert ped.count("7f7fffff") > 50 and ped.count("00000000") > 50 and len(set(ped)) > 100


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="compiled reference not present")
def test_reference_binary_matches_digest_and_product(product_out, built):
    ref = subprocess.run([REF_EXE], capture_output=True, text=True, check=True).stdout
    assert digest(ref)["sha256"] == GOLD["sha256"], "golden is stale: rerun tests/golden/make_velped_golden.py"
    assert ref == product_out
