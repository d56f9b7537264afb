"""Builds libfvs_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m finitevolumesolver_b200.build [--force] [--verbose]

The twelve instantiation units (3 space dimensions x 4 base methods) compile
in parallel; objects are cached under csrc/build/ and rebuilt when a source
they depend on is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libfvs_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: no FMA contraction, results are bit-identical to the
# reference's CPU path (DESIGN.md "Bitwise strategy"); double division and
# sqrt are IEEE-754 correctly rounded by default.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]

# launch shape of the tolerance-arithmetic kernels, measured best (profiles/r01_tuning.md):
# the x sweep holds the next chunk in registers at 3 resident blocks of 128 threads
TOLERANCE_TUNING = ["-DFVS_SWEEP_XS_PREFETCH=2", "-DFVS_SWEEP_XS_MIN_BLOCKS=3"]

HEADERS = ["numerics.cuh", "fastmath.cuh", "launch.h", "sweep_kernels.cuh", "boxes.hpp",
           os.path.join(INCLUDE, "fvs_b200.h")]


def _units():
    units = []
    for d in (1, 2, 3):
        for b in (0, 1, 2, 3):
            units.append(("sweep_inst.cu", f"sweep_D{d}_B{b}.o",
                          [f"-DFVS_INST_D={d}", f"-DFVS_INST_BASE={b}"]))
    # the fused sweep once more with the tolerance arithmetic (fvs_set_arithmetic):
    # the only units with FMA contraction; no build for the exact Riemann solver
    for d in (1, 2, 3):
        for b in (0, 1, 2):
            units.append(("sweep_inst.cu", f"sweep_tol_D{d}_B{b}.o",
                          [f"-DFVS_INST_D={d}", f"-DFVS_INST_BASE={b}", "-DFVS_TOLERANCE_UNIT=1",
                           "-fmad=true"] + TOLERANCE_TUNING))
    units.append(("generic_kernels.cu", "generic_kernels.o", []))
    units.append(("cutcell.cu", "cutcell.o", []))
    units.append(("amr.cu", "amr.o", []))
    units.append(("context.cu", "context.o", []))
    return units


def _newest_header():
    t = 0.0
    for h in HEADERS:
        p = h if os.path.isabs(h) else os.path.join(CSRC, h)
        t = max(t, os.path.getmtime(p))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def _compile(unit, force, verbose, extra):
    src, obj, defs = unit
    src_p = os.path.join(CSRC, src)
    obj_p = os.path.join(OBJ, obj)
    if (not force and os.path.exists(obj_p)
            and os.path.getmtime(obj_p) > max(os.path.getmtime(src_p), _newest_header())):
        return obj_p, ""
    flags = [f for f in NVCC_FLAGS if not (f == "-fmad=false" and "-fmad=true" in defs)]
    cmd = [NVCC] + flags + extra + defs + ["-I", INCLUDE, "-c", src_p, "-o", obj_p]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
    return obj_p, r.stderr


def build(force=False, verbose=False, extra=()):
    os.makedirs(OBJ, exist_ok=True)
    extra = list(extra) + os.environ.get("FVS_NVCC_EXTRA", "").split()
    units = _units()
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda u: _compile(u, force, verbose, list(extra)), units))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    stale = (force or not os.path.exists(LIB)
             or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs))
    if stale:
        cmd = [NVCC, "-shared", "-o", LIB, "-ccbin", "/usr/bin/g++"] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return log


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    if out:
        print(out)
    print(LIB)
