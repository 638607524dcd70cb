"""path_planning_pkg_b200 -- B200-native (sm_100a) local-planner search hot path of path_planning_pkg.

The product is native: `lib/libpp_b200.so` (hand-written CUDA kernels behind the C ABI of
`include/pp_b200.h`) and `lib/libpath_planning_b200.so` (the reference's C++ class API on top of it).
This Python package only holds the build script and a ctypes binding used by tests and bench.py.
Importing it never falls back to a CPU implementation: `_cabi.load()` raises when the CUDA library is
missing, and `pp_create` fails when there is no CUDA device.
"""
from . import _cabi  # noqa: F401
from ._cabi import Context, PPError, make_params, load  # noqa: F401

__all__ = ["Context", "PPError", "make_params", "load"]
