"""csrc/core/pp_gmath.h -- the PP_HD restatement of glibc 2.39's binary32 sinf / cosf / atanf / atan2f / acosf that the EXACT mode
evaluates on the device -- against the libm of this machine (the one the stock reference build, oracle/_ref/libref_oracle.so,
calls).  tests/cpp/gmath_check.cpp compares bit patterns.  Default: every 61st of the 2^32 arguments of the one-argument
functions (+ sincosf == (sinf, cosf), which GCC substitutes in the reference's Dubins.o) and 2*10^7 atan2f pairs, a few seconds;
PP_GMATH_FULL=1 runs all 2^32 arguments and 10^9 pairs (about two minutes on 8 cores; last full run: 0 mismatches)."""
import os
import subprocess

import orc

CHECK = os.path.join(orc.ROOT, "tests", "cpp", "bin", "gmath_check")


def _run(args):
    r = subprocess.run([CHECK] + args, capture_output=True, text=True, timeout=3600)
    rows = {l.split()[0]: (int(l.split()[2]), int(l.split()[4])) for l in r.stdout.splitlines() if " checked " in l}
    return r.returncode, rows, r.stdout


def test_one_argument_functions_bit_identical_to_libm(built):
    full = os.environ.get("PP_GMATH_FULL") == "1"
    rc, rows, out = _run(["exhaustive", "1" if full else "61"])
    assert set(rows) == {"sinf", "cosf", "atanf", "acosf", "sincosf_vs_sinf_cosf", "pp_g_sincosf"}, out
    for name, (n, bad) in rows.items():
        assert n >= (1 << 32) // 61 and bad == 0, out
    assert rc == 0


def test_atan2f_bit_identical_to_libm(built):
    full = os.environ.get("PP_GMATH_FULL") == "1"
    rc, rows, out = _run(["atan2", "1000000000" if full else "20000000", "3"])
    assert rows["atan2f"][0] >= 20000000 and rows["atan2f"][1] == 0, out
    assert rc == 0
