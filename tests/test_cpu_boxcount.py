"""pp_box_count (gather form of the reference's box rasteriser, csrc/core/pp_map.h) against a brute-force scatter with the reference's
forward arithmetic (Grid2D.cpp:127-130): 1 500 random grid headings / box sizes, every cell of a 121 x 121 window."""
import os
import subprocess

import orc

SRC = os.path.join(orc.ROOT, "tests", "cpp", "boxcount_check.cpp")
EXE = os.path.join(orc.ROOT, "tests", "cpp", "bin", "boxcount_check")


def test_gather_counts_equal_scatter_counts():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++14", "-o", EXE, SRC])
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-1000:]
