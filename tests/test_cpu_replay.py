"""SURVEY.md §8(f) N4, CPU side of the replay harness: the reference's UNMODIFIED ROS node src/local_planner.cpp, compiled
against the in-process ROS stand-in (tests/ros_stubs/ros/ros.h) and the UNMODIFIED reference library, replays a scripted
timeline of odometry / waypoint / object / lane messages (tests/replay_scenario.py) and publishes trajectories.  These
tests pin the committed golden (tests/golden/replay_ref.json) to that binary; tests/test_gpu_replay.py runs the same node on
libpath_planning_b200.so and demands the same trajectories."""
import hashlib
import json
import os
import sys

import pytest

import orc
import replay_scenario as R

sys.path.insert(0, os.path.join(orc.ROOT, "tests", "golden"))
from make_replay_golden import SEEDS, digest, run_node  # noqa: E402

GOLD = json.load(open(os.path.join(orc.ROOT, "tests", "golden", "replay_ref.json")))
REF_NODE = os.path.join(orc.ROOT, "oracle", "_ref", "local_planner_ref")


def test_scripts_are_reproducible():
    for seed in SEEDS:
        assert hashlib.sha256(R.make_script(seed).encode()).hexdigest() == GOLD[str(seed)]["script_sha256"]


def test_golden_covers_the_callers_paths():
    """Every tick after the first waypoint publishes one trajectory of (x, y, heading, velocity) rows."""
    for seed in SEEDS:
        g = GOLD[str(seed)]["ref"]
        assert g["ticks"] == 18 and len(g["pubs"]) == 17
        assert [p["tick"] for p in g["pubs"]] == list(range(1, 18))
        assert all(p["topic"] == "/local_planner/trajectory" and p["n"] % 4 == 0 and p["n"] >= 8 for p in g["pubs"])


@pytest.mark.skipif(not os.path.exists(REF_NODE), reason="compiled reference node not present")
@pytest.mark.parametrize("seed", SEEDS)
def test_reference_node_reproduces_golden(seed, built):
    pubs, ticks, times, log = run_node(REF_NODE, R.make_script(seed))
    assert ticks == GOLD[str(seed)]["ref"]["ticks"]
    assert digest(pubs) == GOLD[str(seed)]["ref"]["pubs"], "golden is stale: rerun tests/golden/make_replay_golden.py"
    assert "Velocity Generator: Failed" not in log


def test_product_node_fails_loudly_without_a_gpu(tmp_path):
    """The same node on libpath_planning_b200.so has no CPU path: on a machine without a CUDA device the planner's
    constructor aborts with the C ABI's error text instead of computing anything on the host."""
    import subprocess
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    node = os.path.join(orc.ROOT, "tests", "cpp", "bin", "local_planner_b200")
    if not os.path.exists(node):
        pytest.skip("local_planner_b200 is built from /root/reference/src/local_planner.cpp, absent here")
    script = tmp_path / "script.txt"
    script.write_text(R.make_script(0))
    r = subprocess.run([node], capture_output=True, text=True, env=dict(os.environ, PP_REPLAY_SCRIPT=str(script), PP_REPLAY_OUT=str(tmp_path / "out.txt")))
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr
    assert not (tmp_path / "out.txt").exists() or "pub" not in (tmp_path / "out.txt").read_text()
