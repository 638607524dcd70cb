import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (if stale) the native pieces the CPU tests need: host emulation + oracles."""
    from path_planning_pkg_b200 import build
    build.build_host_emul(verbose=False)
    build.build_gmath_check(verbose=False)
    build.build_oracle(verbose=False)
    return True
