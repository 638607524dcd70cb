"""The round-1 math policy stays selectable as a build variant: lib/libpp_b200_pinned.so (-DPP_MATH_PINNED, float transcendentals
evaluated in double and rounded once, libm-version independent) must still equal the reference objects linked against the same
definition (oracle/_ref/libref_oracle_crm.so).  The default library restates glibc and is checked against the stock build
everywhere else.  Runs in a subprocess because the ctypes binding loads one library per process (PP_B200_LIB)."""
import os
import subprocess
import sys

import pytest

import orc

pytestmark = pytest.mark.gpu
VARIANT = os.path.join(orc.ROOT, "path_planning_pkg_b200", "lib", "libpp_b200_pinned.so")

SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import orc, path_planning_pkg_b200 as pp
P = orc.ref_test_params()
ctx = pp.Context(pp._cabi.params_from(P), num_groups=1, device=0)
crm = orc.crm(P)
for o in (ctx, crm):
    orc.setup_ref_test_scenario(o)
a = ctx.find_path(2.0, orc.REF_TEST_START); b = crm.find_path(2.0, orc.REF_TEST_START)
assert a["n_pops"] == b["n_pops"] and a["cost"] == b["cost"]
for f in a["pops"].dtype.names:
    assert np.array_equal(a["pops"][f].view(np.uint32), b["pops"][f].view(np.uint32)), f
assert np.array_equal(a["path"].view(np.uint32), b["path"].view(np.uint32))
rs = np.random.RandomState(2)
xyh = np.stack([rs.uniform(2, 28, 5000), rs.uniform(2, 28, 5000), rs.uniform(-3.14, 3.14, 5000)], 1).astype(np.float32)
assert np.array_equal(ctx.apf(xyh).view(np.uint32), crm.apf(xyh).view(np.uint32))
print("pinned variant ok", a["n_pops"])
'''


@pytest.mark.skipif(not (os.path.exists(VARIANT) and os.path.exists(orc.CRM_SO)), reason="pinned build variant / crm oracle not built")
def test_pinned_variant_equals_pinned_reference():
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=orc.ROOT, env=dict(os.environ, PP_B200_LIB=VARIANT),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "pinned variant ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
