"""Tuning helper: builds libfvs_b200_<tag>.so with extra -D flags for the
3-D HLLEM-Larrouturou unit only (the bench configuration); the other objects
are reused from the normal build.  Select with FVS_B200_LIB=<path>.

    python tools/build_variant.py <tag> [-DNAME=VALUE ...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finitevolumesolver_b200 import build as b  # noqa: E402


def compile_unit(tag, flags, tol):
    obj = os.path.join(b.OBJ, f"variant_{tag}_{'tol' if tol else 'ieee'}.o")
    base_flags = list(b.NVCC_FLAGS)
    if tol:
        base_flags = [f for f in base_flags if f != "-fmad=false"] + ["-fmad=true",
                                                                      "-DFVS_TOLERANCE_UNIT=1"]
    cmd = [b.NVCC] + base_flags + flags + ["-DFVS_INST_D=3", "-DFVS_INST_BASE=0", "-DFVS_INST_MINIMAL", "-Xptxas", "-v",
           "-I", b.INCLUDE, "-c", os.path.join(b.CSRC, "sweep_inst.cu"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit("\n".join(x for x in r.stderr.splitlines() if "error" in x.lower()))
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if "fused_sweep" in l and "ELi1ELi2E" in l and "Function properties" in l:
            print(tag, "tol " if tol else "ieee", l.split()[-1][:48], "|", lines[i + 1].strip(), "|",
                  lines[i + 2].strip())
    return obj


def main():
    # --tol: vary the tolerance-arithmetic unit, --both: both units, default: the IEEE unit
    tag, flags = sys.argv[1], sys.argv[2:]
    which = [True] if "--tol" in flags else [False, True] if "--both" in flags else [False]
    flags = [f for f in flags if f not in ("--tol", "--both")]
    b.build()
    replaced = {True: "sweep_tol_D3_B0.o", False: "sweep_D3_B0.o"}
    objs = [os.path.join(b.OBJ, u[1]) for u in b._units()
            if u[1] not in [replaced[w] for w in which]]
    objs += [compile_unit(tag, flags, w) for w in which]
    lib = os.path.join(b.HERE, f"libfvs_b200_{tag}.so")
    subprocess.check_call([b.NVCC, "-shared", "-o", lib, "-ccbin", "/usr/bin/g++"] + objs + ["-ldl"])
    print(lib)


if __name__ == "__main__":
    main()
