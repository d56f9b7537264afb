"""Builds the C++ test binaries under tests/cpp/ (g++ only, no CUDA needed).

test_shim links the product library (libfvs_b200.so) and the checker
(liboracle.so); the binaries are git-ignored but travel to the GPU box.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "finitevolumesolver_b200")
ORACLE = os.path.join(ROOT, "oracle")
CXX = "/usr/bin/g++"


def build(force=False):
    out = os.path.join(HERE, "test_shim")
    src = os.path.join(HERE, "test_shim.cpp")
    deps = [src, os.path.join(ROOT, "include", "fub_b200", "HyperbolicMethod.hpp"),
            os.path.join(ROOT, "include", "fvs_b200.h")]
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) > max(os.path.getmtime(d) for d in deps)):
        return out
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
           src, "-o", out, "-L", PKG, "-lfvs_b200", "-L", ORACLE, "-loracle",
           f"-Wl,-rpath,$ORIGIN/../../finitevolumesolver_b200", f"-Wl,-rpath,$ORIGIN/../../oracle",
           "-Wl,--allow-shlib-undefined"]
    subprocess.check_call(cmd)
    return out


def build_fastmath_check(force=False):
    """tests/cuda/fastmath_check: FastMath vs the IEEE operators bit for bit,
    ToleranceMath's error bound (nvcc cross-compiles without a GPU)."""
    cuda = os.path.join(os.path.dirname(HERE), "cuda")
    out = os.path.join(cuda, "fastmath_check")
    src = os.path.join(cuda, "fastmath_check.cu")
    deps = [src, os.path.join(PKG, "csrc", "fastmath.cuh"), os.path.join(cuda, "fastmath_cases.h")]
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) > max(os.path.getmtime(d) for d in deps)):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-fmad=false", "-ccbin", CXX, "-o", out, src])
    return out


def build_fastmath_host(force=False):
    """tests/cpp/fastmath_host: csrc/fastmath.cuh compiled for the CPU (hardware FMA,
    emulated MUFU seeds) against `/` and sqrt()."""
    out = os.path.join(HERE, "fastmath_host")
    src = os.path.join(HERE, "fastmath_host.cpp")
    deps = [src, os.path.join(PKG, "csrc", "fastmath.cuh"),
            os.path.join(os.path.dirname(HERE), "cuda", "fastmath_cases.h")]
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) > max(os.path.getmtime(d) for d in deps)):
        return out
    subprocess.check_call([CXX, "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-o", out, src])
    return out


if __name__ == "__main__":
    print(build_fastmath_host(force=True))
    print(build_fastmath_check(force=True))
    print(build(force=True))
