"""The C++17 policy shim (include/fub_b200/HyperbolicMethod.hpp): compiles
against the C ABI on CPU; on a GPU its integrator reproduces the oracle bit for
bit (tests/cpp/test_shim.cpp)."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))


@pytest.fixture(scope="module")
def shim_binary(oracle):
    from finitevolumesolver_b200 import capi
    capi.lib()
    import build_cpp_tests
    return build_cpp_tests.build()


def test_shim_compiles_and_refuses_to_run_without_gpu(shim_binary):
    from finitevolumesolver_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("GPU present; covered by test_shim_matches_oracle")
    r = subprocess.run([shim_binary], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, r.stdout + r.stderr  # no CPU fallback
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_shim_matches_oracle(shim_binary):
    r = subprocess.run([shim_binary], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all shim checks passed" in r.stdout
