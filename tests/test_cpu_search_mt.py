"""CPU race check of the search cores: tests/cpp/search_mt.cpp runs the product's pp_search_kpop on 4 and 8 host threads and
pp_search_exact on 8 and 32 (one thread per lane, pthread barriers for the CTA / warp barriers, GCC atomics for the device
atomics) under ThreadSanitizer and compares every result with the single-lane run of the same code, bit for bit.  (compute-sanitizer's racecheck is not
available on the GPU pool, so this is where unordered shared-memory accesses would show up.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
import scenarios as S

BIN = os.path.join(orc.ROOT, "tests", "cpp", "bin")
SRC = os.path.join(orc.ROOT, "tests", "cpp", "search_mt.cpp")


@pytest.fixture(scope="module", params=[0, 1], ids=["sequential-commit", "speculative-walks"])
def exe(request):
    """search_mt built twice: the default EXACT core and the -DPP_EXACT_SPEC=1 variant (speculative parallel walks)."""
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(BIN, "search_mt" + ("_spec" if request.param else ""))
    cmd = ["g++", "-std=c++14", "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-Wno-unknown-pragmas",
           f"-DPP_EXACT_SPEC={request.param}", "-o", out, SRC, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return out


def _write_scenario(path, P, goal, start, boxes, conf, radius, grid_map, h1, queries, k):
    with open(path, "wb") as f:
        f.write(bytes(P))
        f.write(np.asarray(goal, np.float32).tobytes()); f.write(np.asarray(start, np.float32).tobytes())
        f.write(np.int32(len(boxes)).tobytes())
        f.write(np.ascontiguousarray(boxes, np.float32).tobytes()); f.write(np.ascontiguousarray(conf, np.float32).tobytes())
        f.write(np.float32(radius).tobytes())
        f.write(np.ascontiguousarray(grid_map, np.float32).tobytes()); f.write(np.ascontiguousarray(h1, np.float32).tobytes())
        f.write(np.int32(len(queries)).tobytes()); f.write(np.ascontiguousarray(queries, np.float32).tobytes())
        f.write(np.int32(k).tobytes())


def _field(port):
    d = orc.field2d(port)
    return np.where(d >= 0, d, 3.0e38).astype(np.float32)


@pytest.mark.parametrize("k", [32, 5])
def test_multilane_race_free_and_identical(exe, tmp_path, k):
    runs = []
    # the reference's own test scenario (lane lines + boxes)
    P = orc.ref_test_params()
    port = orc.port(P)
    orc.setup_ref_test_scenario(port)
    s = orc.REF_TEST_START
    runs.append((P, orc.REF_TEST_GOAL, orc.REF_TEST_START, orc.REF_TEST_BOXES, np.full(3, 0.75, np.float32), 2.5, port,
                 np.array([[s[0], s[1], s[2], 2.0], [s[0] + 1.0, s[1] - 0.5, s[2] + 0.2, 0.5]], np.float32)))
    # a C1 clutter scenario
    sc = S.c1_scenario(4)
    P2 = orc.make_params(grid_size=sc["grid_size"], resolution=sc["resolution"])
    port2 = orc.port(P2)
    S.build_map(port2, sc)
    runs.append((P2, sc["goal"], sc["frame_start"], sc["boxes"], sc["conf"], S.APF_ADDED_RADIUS, port2, sc["queries"][:1]))
    for i, (prm, goal, start, boxes, conf, radius, o, queries) in enumerate(runs):
        path = str(tmp_path / f"scenario{i}.bin")
        _write_scenario(path, prm, goal, start, boxes, conf, radius, o.get_map(), _field(o), queries, k)
        env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66 report_signal_unsafe=0")
        r = subprocess.run([exe, path], capture_output=True, text=True, env=env, timeout=900)
        print(r.stdout)
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:3000]
        assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
        assert "MISMATCH" not in r.stdout and "identical" in r.stdout
        assert "EXACT on the carried cache" in r.stdout          # the planner-object-history pass (incl. the stamp wrap) ran
