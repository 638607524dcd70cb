"""The instruction footprint of pp_search_kernel is a performance property of the product (DESIGN.md sections 3, 7: at bench
occupancy the kernel is bound by instruction fetch; cutting its SASS from 189 KB to 123 KB bought +43 % throughput), and it is easy
to lose without noticing: one loop over a small register array and the compiler unrolls the 2D A*'s neighbour body eight times
again (46 KB for that function alone).  This test disassembles the object build() produced and holds the line.  CPU only: nvcc
cross-compiles sm_100a here, cuobjdump / nvdisasm read the cubin."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "path_planning_pkg_b200", "lib", "obj", "libpp_b200_cabi.o")

pytestmark = pytest.mark.skipif(not (shutil.which("cuobjdump") and shutil.which("nvdisasm")), reason="needs the CUDA binary utilities")


@pytest.fixture(scope="module")
def footprint():
    sys.path.insert(0, ROOT)
    from path_planning_pkg_b200 import build
    build.build_cuda(force=not os.path.exists(OBJ), verbose=False)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_footprint.py"), "pp_search_kernel", OBJ],
                         capture_output=True, text=True, check=True).stdout
    total = int(out.split(":")[1].split("bytes")[0])
    per_fn = {}
    for ln in out.splitlines()[1:]:
        parts = ln.split()
        if len(parts) >= 4 and parts[0].isdigit():
            per_fn[parts[3].split(":")[-1]] = per_fn.get(parts[3].split(":")[-1], 0) + int(parts[0])
    return total, per_fn, out


def test_search_kernel_footprint_holds(footprint):
    total, per_fn, out = footprint
    print(out.splitlines()[0])
    assert total <= 135 * 1024, f"pp_search_kernel grew to {total} bytes of SASS (123 KB when this test was written)\n{out}"


def test_lazy_astar_neighbour_loop_is_not_unrolled(footprint):
    _, per_fn, out = footprint
    lazy = per_fn.get("pp_lazy_astar", 0)
    assert 0 < lazy <= 10 * 1024, f"pp_lazy_astar is {lazy} bytes of SASS (7 KB rolled, 46 KB when the neighbour loop was unrolled)\n{out}"


def test_mirrored_tree_code_is_shared(footprint):
    _, per_fn, out = footprint
    assert per_fn.get("pp_rb_erase", 0) <= 8 * 1024 and per_fn.get("pp_rb_insert_and_rebalance", 0) <= 4 * 1024, out
