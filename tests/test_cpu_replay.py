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
REF_NODE = os.path.join(orc.ROOT, "oracle", "_ref", "local_planner_ref_crm")


def test_scripts_are_reproducible():
    for seed in SEEDS:
        assert hashlib.sha256(R.make_script(seed).encode()).hexdigest() == GOLD[str(seed)]["script_sha256"]


def test_golden_covers_the_callers_paths():
    """Every tick after the first waypoint publishes one trajectory of (x, y, heading, velocity) rows."""
    for seed in SEEDS:
        g = GOLD[str(seed)]["ref_crm"]
        assert g["ticks"] == 18 and len(g["pubs"]) == 17
        assert [p["tick"] for p in g["pubs"]] == list(range(1, 18))
        assert all(p["topic"] == "/local_planner/trajectory" and p["n"] % 4 == 0 and p["n"] >= 8 for p in g["pubs"])


@pytest.mark.skipif(not os.path.exists(REF_NODE), reason="compiled reference node not present")
@pytest.mark.parametrize("seed", SEEDS)
def test_reference_node_reproduces_golden(seed, built):
    pubs, ticks, times, log = run_node(REF_NODE, R.make_script(seed))
    assert ticks == GOLD[str(seed)]["ref_crm"]["ticks"]
    assert digest(pubs) == GOLD[str(seed)]["ref_crm"]["pubs"], "golden is stale: rerun tests/golden/make_replay_golden.py"
    assert "Velocity Generator: Failed" not in log
