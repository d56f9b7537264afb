"""csrc/fastmath.cuh on the device: the branch-free division / square root of
the default arithmetic equal `/` and sqrt() bit for bit wherever their validity
flag is clean (random bit patterns, moderate operands, special values), and the
opt-in ToleranceMath stays within 2 ulp of them (tests/cuda/fastmath_check.cu)."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))

pytestmark = pytest.mark.gpu


def test_fastmath_bitwise_and_tolerance_bound():
    import build_cpp_tests
    binary = build_cpp_tests.build_fastmath_check()
    r = subprocess.run([binary, "296"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    fields = r.stdout.split()
    vals = dict(zip(fields[0::2], map(int, fields[1::2])))
    assert vals["mismatch_div"] == 0 and vals["mismatch_sqrt"] == 0 and vals["mismatch_divz"] == 0
    assert vals["zero_divz"] > 0.2 * vals["tested"]
    assert vals["tested"] > 1e9
    assert vals["tol_div_ulp"] <= 2 and vals["tol_sqrt_ulp"] <= 2, vals
