"""Generates the committed golden fixtures (run HERE, where /root/reference exists; the GPU box only reads the .npy/.npz).

(1) The three result vectors the reference's author pasted into its plotting scripts:
      utils/hybrid_astar/plot.py:47-51   43-point Hybrid A* path          -> hybrid_astar_path.npy
      utils/dubins_paths.py:6            73-point RSL Dubins path          -> dubins_rsl_path.npy
      utils/vehicle_mode.py:12           33-point simulate_action roll-out -> vehicle_rollout.npy
    They are parsed out of the reference files (values only, 6 significant digits).
(2) Outputs of the unmodified reference itself, run through oracle/_ref (both libm flavours):
      golden_search.npz   pop sequence / path / cost of the reference's own scenario
      golden_c1.npz       cost, pops, path checksums for C1 seeds 0..15
      golden_map_c2.npz   CRC of the 2048^2 map after each of the 3 C2 rounds + sparse sample
"""
import os
import re
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402
import scenarios as S  # noqa: E402

REF = "/root/reference"


def _parse_array(path, lineno_from, lineno_to):
    lines = open(path).read().split("\n")[lineno_from - 1:lineno_to]
    txt = " ".join(lines)
    txt = txt[txt.index("np.array(") + len("np.array("):]
    depth, end = 0, None
    for i, ch in enumerate(txt):
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
            if depth == 0:
                end = i + 1
                break
    rows = re.findall(r"\[([^\[\]]+)\]", txt[:end])
    return np.array([[float(v) for v in r.split(",")] for r in rows], np.float64)


def main():
    a = _parse_array(f"{REF}/utils/hybrid_astar/plot.py", 47, 51)
    assert a.shape == (43, 3), a.shape
    np.save(f"{HERE}/hybrid_astar_path.npy", a)
    b = _parse_array(f"{REF}/utils/dubins_paths.py", 6, 6)
    assert b.shape == (73, 3), b.shape
    np.save(f"{HERE}/dubins_rsl_path.npy", b)
    c = _parse_array(f"{REF}/utils/vehicle_mode.py", 12, 12)
    assert c.shape == (33, 2), c.shape
    np.save(f"{HERE}/vehicle_rollout.npy", c)

    # reference's own scenario through the compiled reference
    out = {}
    for name, mk in (("ref", orc.ref), ("crm", orc.crm)):
        P = orc.ref_test_params()
        o = mk(P)
        orc.setup_ref_test_scenario(o)
        r = o.find_path(2.0, orc.REF_TEST_START)
        out[f"{name}_pops"] = r["pops"]
        out[f"{name}_path"] = r["path"]
        out[f"{name}_curv"] = r["curvature"]
        out[f"{name}_cost"] = np.float32(r["cost"])
        out[f"{name}_map"] = o.get_map()
    np.savez_compressed(f"{HERE}/golden_search.npz", **out)

    rows = []
    for seed in range(16):
        sc = S.c1_scenario(seed)
        P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
        rec = [seed]
        for mk in (orc.ref, orc.crm):
            o = mk(P)
            S.build_map(o, sc)
            q = sc["queries"][0]
            r = o.find_path(float(q[3]), q[:3])
            rec += [int(r["success"]), float(r["cost"]), r["n_pops"], len(r["path"]),
                    zlib.crc32(r["pops"][["ci", "cj", "bin"]].tobytes()), zlib.crc32(r["path"].tobytes()),
                    zlib.crc32(o.get_map().tobytes())]
        rows.append(rec)
    np.savez_compressed(f"{HERE}/golden_c1.npz", rows=np.array(rows, np.float64))

    sc = S.c2_scenario()
    P = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    o = orc.ref(P)
    o.update_goal(sc["goal"], sc["frame_start"])
    crcs, samples = [], []
    for _ in range(sc["rounds"]):
        o.update_boxes_2d(sc["boxes"], sc["conf"])
        o.decay()
        m = o.get_map()
        crcs.append(zlib.crc32(m.tobytes()))
        samples.append(m[::64, ::64].copy())
    thr = o.consts().log_threshold
    np.savez_compressed(f"{HERE}/golden_map_c2.npz", crcs=np.array(crcs, np.uint64), samples=np.array(samples),
                        occupied=np.int64((m >= thr).sum()))
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
