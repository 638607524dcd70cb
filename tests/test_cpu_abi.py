"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/pp_b200.h declares,
its structs have the layout the host side assumes, and -- with no GPU -- it refuses to create a context instead of
falling back to a CPU path.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import orc

ROOT = orc.ROOT


@pytest.fixture(scope="module")
def lib():
    from path_planning_pkg_b200 import build
    build.build_cuda(verbose=False)
    import path_planning_pkg_b200 as pp
    return pp.load()


def test_library_exports_every_declared_symbol(lib):
    import path_planning_pkg_b200 as pp
    declared = pp._cabi.exported_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_header_cites_reference_interfaces():
    txt = open(os.path.join(ROOT, "include", "pp_b200.h")).read()
    for ref in ("lib/HybridAStar.cpp:68-88", "lib/Grid2D.cpp:99-139", "Grid2D.cpp:197-208", "lib/Grid3D.cpp:47-74",
                "lib/Dubins.cpp:19-69", "lib/VehicleModel.cpp:63-105", "lib/AStar.cpp:100-113"):
        assert ref in txt, ref


def test_struct_layouts_match_native(lib):
    import path_planning_pkg_b200 as pp
    # sizes the native side static_asserts / memcpy's on
    assert C.sizeof(pp._cabi.Params) == C.sizeof(orc.Params) == 19 * 4 + 2 * 16 * 4
    assert pp._cabi.STATE_DT.itemsize == 40 and pp._cabi.POP_DT.itemsize == 32
    assert pp._cabi.QUERY_DT.itemsize == 20 and pp._cabi.RESULT_DT.itemsize == 48
    assert C.sizeof(pp._cabi.SearchOpts) == 32


def test_no_cpu_fallback(lib):
    """Without a CUDA device pp_create must fail loudly (PP_ERR_NO_DEVICE); with one this test is skipped."""
    import path_planning_pkg_b200 as pp
    if lib.pp_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(pp.PPError) as e:
        pp.Context(pp.make_params())
    assert "no CUDA device" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """Nothing under the product package or include/ may import, include or link anything under oracle/."""
    bad = []
    for base in ("path_planning_pkg_b200", "include"):
        for r, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp", ".c")):
                    txt = open(os.path.join(r, f), errors="ignore").read()
                    code = "\n".join(l for l in txt.split("\n") if not l.lstrip().startswith(("//", "*", "/*", "#  ", '"""')))
                    if re.search(r'#include\s+"[^"]*oracle|import\s+orc\b|CDLL\([^)]*oracle|dlopen\([^)]*oracle', code):
                        bad.append(os.path.join(r, f))
    # build.py only *builds* the checker (make -C oracle); it never loads it
    bad = [b for b in bad if not b.endswith("build.py")]
    assert not bad, bad


def test_cuda_library_has_sm100a_code(lib):
    import path_planning_pkg_b200 as pp
    out = subprocess.run(["cuobjdump", "-lelf", pp._cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
