"""csrc/fastmath.cuh on the CPU (no GPU needed): the branch-free division --
incl. the zero-tolerant divz over the positive divisor window and div_any -- and
the square root equal `/` and sqrt() bit for bit wherever their validity flag is
clean.  The refinement, the negated-remainder handling of +-0 numerators and the
validity tests do not depend on the seed, which is emulated to MUFU accuracy
(tests/cuda/fastmath_cases.h; tests/test_gpu_fastmath.py runs the same cases on
the device with the real seeds)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))


def test_fastmath_host_bitwise():
    import build_cpp_tests
    binary = build_cpp_tests.build_fastmath_host()
    r = subprocess.run([binary, "40000000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    fields = r.stdout.split()
    vals = dict(zip(fields[0::2], map(int, fields[1::2])))
    assert vals["mismatch_div"] == 0 and vals["mismatch_sqrt"] == 0 and vals["mismatch_divz"] == 0
    assert vals["tested"] == 40000000
    # deterministic sweep over the exponents around the thresholds of the folded tests
    assert vals["threshold_sweep_mismatch"] == 0
    assert vals["threshold_sweep_clean"] > 0.1 * vals["threshold_sweep_tested"]
    # the zero numerators are accepted by the fast path, not sent to the fallback
    assert vals["zero_divz"] > 0.2 * vals["tested"]
    # more than a third of the zero-tolerant divisions ran inside the window
    assert vals["flagged_divz"] < 0.65 * vals["tested"]
    assert vals["tol_div_ulp"] <= 2 and vals["tol_sqrt_ulp"] <= 2, vals
