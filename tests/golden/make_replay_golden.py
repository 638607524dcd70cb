"""Generates tests/golden/replay_ref.json (run HERE, where /root/reference exists): the trajectories the reference's
UNMODIFIED ROS node (src/local_planner.cpp) publishes when it runs on the UNMODIFIED reference library behind the
in-process ROS stand-in (tests/ros_stubs), for the scripts of tests/replay_scenario.py.  Built by oracle/Makefile as
oracle/_ref/local_planner_ref (stock glibc libm) and local_planner_ref_crm (pinned libm = the device's definition of the
float transcendentals).  Stored per published message: tick, sample count, SHA-256 of the data words."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
import replay_scenario as R  # noqa: E402

SEEDS = (0, 1, 2)


def run_node(exe, script_text, env_extra=None, timeout=600):
    """Runs a local_planner binary on a script; returns (published messages, ticks, per-tick milliseconds, stdout)."""
    with tempfile.TemporaryDirectory() as d:
        sp, op, tp = os.path.join(d, "script.txt"), os.path.join(d, "out.txt"), os.path.join(d, "times.txt")
        open(sp, "w").write(script_text)
        env = dict(os.environ, PP_REPLAY_SCRIPT=sp, PP_REPLAY_OUT=op, PP_REPLAY_TIMES=tp)
        env.update(env_extra or {})
        r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=timeout)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        pubs, ticks = R.parse_output(open(op).read())
        times = [float(l.split()[1]) for l in open(tp).read().split("\n") if l.strip()]
        return pubs, ticks, times, r.stdout


def digest(pubs):
    return [{"tick": t, "topic": topic, "n": n, "sha256": hashlib.sha256(" ".join(words).encode()).hexdigest()} for t, topic, n, words in pubs]


if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    gold = {}
    for seed in SEEDS:
        script = R.make_script(seed)
        entry = {"script_sha256": hashlib.sha256(script.encode()).hexdigest()}
        for flavour in ("ref", "ref_crm"):
            pubs, ticks, times, _ = run_node(os.path.join(ROOT, "oracle", "_ref", "local_planner_" + flavour), script)
            entry[flavour] = {"ticks": ticks, "pubs": digest(pubs)}
        gold[str(seed)] = entry
        same = sum(1 for a, b in zip(entry["ref"]["pubs"], entry["ref_crm"]["pubs"]) if a == b)
        print(f"seed {seed}: {len(entry['ref_crm']['pubs'])} trajectories, stock == pinned libm in {same}")
    json.dump(gold, open(os.path.join(HERE, "replay_ref.json"), "w"), indent=1)
