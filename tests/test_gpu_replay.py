"""SURVEY.md §8(f) N4 on the GPU: the reference's UNMODIFIED ROS node (src/local_planner.cpp) linked against THIS repo's
headers and libpath_planning_b200.so (tests/cpp/bin/local_planner_b200, built by path_planning_pkg_b200.build where the
reference tree exists) replays scripted sessions behind the in-process ROS stand-in (tests/ros_stubs).  Every trajectory it
publishes -- path from HybridAStar<float>::find_path on the device (EXACT mode, planner-object history), velocity profile
from VelocityGenerator, pedestrian cap from PedestrianHandler -- must equal, word for word, what the same node publishes on
the unmodified reference library with the stock glibc (golden: tests/golden/replay_ref.json; live when oracle/_ref has it)."""
import json
import os
import sys

import pytest

import orc
import replay_scenario as R

sys.path.insert(0, os.path.join(orc.ROOT, "tests", "golden"))
from make_replay_golden import SEEDS, digest, run_node  # noqa: E402

pytestmark = pytest.mark.gpu
NODE = os.path.join(orc.ROOT, "tests", "cpp", "bin", "local_planner_b200")
REF_NODE = os.path.join(orc.ROOT, "oracle", "_ref", "local_planner_ref")
GOLD = json.load(open(os.path.join(orc.ROOT, "tests", "golden", "replay_ref.json")))


@pytest.mark.skipif(not os.path.exists(NODE), reason="local_planner_b200 is built from /root/reference/src/local_planner.cpp, absent here")
@pytest.mark.parametrize("seed", SEEDS)
def test_unmodified_node_publishes_the_reference_trajectories(seed):
    script = R.make_script(seed)
    pubs, ticks, times, log = run_node(NODE, script)
    gold = GOLD[str(seed)]["ref"]
    assert ticks == gold["ticks"] and len(pubs) == len(gold["pubs"])
    mine = digest(pubs)
    bad = [a["tick"] for a, b in zip(mine, gold["pubs"]) if a != b]
    detail = ""
    if bad and os.path.exists(REF_NODE):
        rp, _, _, _ = run_node(REF_NODE, script)
        k = bad[0] - 1
        diff = [i for i, (x, y) in enumerate(zip(pubs[k][3], rp[k][3])) if x != y]
        detail = f"; tick {bad[0]}: n {pubs[k][2]} vs {rp[k][2]}, first differing words {diff[:8]}"
    assert not bad, f"trajectories differ from the reference node at ticks {bad}{detail}"
    assert "Hybrid A*: Failed" not in log and "Velocity Generator: Failed" not in log
    out_dir = os.path.join(orc.ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        ref_ms = run_node(REF_NODE, script)[2] if os.path.exists(REF_NODE) else None
        json.dump({"seed": seed, "b200_ms_per_tick": times, "reference_cpu_ms_per_tick": ref_ms},
                  open(os.path.join(out_dir, f"replay_times_seed{seed}.json"), "w"))
