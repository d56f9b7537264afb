"""Generates the regression fixtures under tests/golden/ from the CPU oracle.

The reference (C++ with AMReX / Eigen / Boost / Vc) can be neither compiled
nor imported in this environment, so these are NOT outputs of the reference
binary: they pin the oracle restatement (which is itself pinned against the
reference's unit tests, tests/test_oracle_golden.py) against accidental
change, and give the GPU tests a checker that does not need the oracle built.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import (complete_from_prim, eb_fab_type, eb_geometry, polygon_phi,  # noqa: E402
                     riemann_state)
from oracle import oracle as orc  # noqa: E402


def sod_1d():
    """200 cells on [-1, 1], reflective ends, gcw 2/0, Godunov
    splitting, MUSCL-Hancock<HLLEM, VanLeer>... (examples/AMReX/PerfectGas1d.cpp:70-131)."""
    n = 200
    dx = [2.0 / n]
    lvl = orc.Level(1, [0], [n - 1], dx, [([0], [n - 1])], scratch_gcw=2, flux_gcw=0)
    lvl.set_method(orc.BASE_HLLEM_LARROUTUROU, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_VANLEER)
    lvl.add_bc(0, 0, orc.BC_REFLECTIVE)
    lvl.add_bc(0, 1, orc.BC_REFLECTIVE)
    arr, lo, hi = lvl.scratch(0)
    x = -1.0 + (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    rho = np.where(x < 0.0, 1.0, 0.125)[None, None, :]
    p = np.where(x < 0.0, 1.0, 0.1)[None, None, :]
    arr[...] = complete_from_prim(1, rho, [np.zeros_like(rho)], p)
    init = arr.copy()
    lvl.fill_ghost_layer()
    dts, _ = lvl.run(0.5, 0.25, -1, orc.SPLIT_GODUNOV)
    return {"init": init, "final": arr.copy(), "dts": np.array(dts)}


def riemann_3d():
    n = 20
    dx = [2.0 / n] * 3
    lvl = orc.Level(3, [0] * 3, [n - 1] * 3, dx, [([0] * 3, [n - 1] * 3)], scratch_gcw=8, flux_gcw=6)
    lvl.set_method(orc.BASE_HLLEM_LARROUTUROU, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_VANLEER)
    for d in range(3):
        for s in (0, 1):
            lvl.add_bc(d, s, orc.BC_TRANSMISSIVE)
    arr, lo, hi = lvl.scratch(0)
    arr[...] = riemann_state(3, lo, hi, dx, [-1.0] * 3)
    lvl.fill_ghost_layer()
    dts, _ = lvl.run(0.5, 1e9, 8, orc.SPLIT_STRANG)
    g = 8
    return {"final_valid": arr[:, g:-g, g:-g, g:-g].copy(), "dts": np.array(dts)}


def wedge_2d():
    n = (40, 28)
    dx = [0.15 / n[0], 0.10 / n[1]]
    lvl = orc.Level(2, [0, 0, 0], [n[0] - 1, n[1] - 1, 0], dx, [([0, 0, 0], [n[0] - 1, n[1] - 1, 0])],
                    scratch_gcw=8, flux_gcw=6)
    lvl.set_method(orc.BASE_HLLEM_LARROUTUROU, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_VANLEER)
    lvl.enable_eb()
    arr, lo, hi = lvl.scratch(0)
    g = eb_geometry(2, lo, hi, dx + [1.0], [0, 0, 0],
                    polygon_phi([(0.02, 0.05), (0.05, 0.0655), (0.05, 0.0345)]))
    lvl.set_eb(0, g, eb_fab_type(g))
    for d in range(2):
        for s in (0, 1):
            lvl.add_bc(d, s, orc.BC_TRANSMISSIVE)
    x = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    rho = np.ones(arr.shape[1:])
    p = np.where(x < 0.015, 5 * 101325.0, 101325.0)[None, None, :] + np.zeros_like(rho)
    arr[...] = complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, p) * (g["alpha"][0] > 0)
    lvl.fill_ghost_layer()
    dts = []
    for _ in range(10):
        dt = 0.4 * min(lvl.compute_stable_dt(0), lvl.compute_stable_dt(1))
        ok, _, _ = lvl.advance_hierarchy(dt, orc.SPLIT_STRANG)
        assert ok
        dts.append(dt)
    return {"final": arr.copy(), "dts": np.array(dts), "is_cut_cell": lvl.is_cut_cell(0),
            "alpha": g["alpha"], "beta0": g["beta"][0], "beta1": g["beta"][1], "bcent": g["bcent"],
            "bnorm": g["bnorm"]}


if __name__ == "__main__":
    orc.build()
    np.savez_compressed(os.path.join(HERE, "sod_1d.npz"), **sod_1d())
    np.savez_compressed(os.path.join(HERE, "riemann_3d.npz"), **riemann_3d())
    np.savez_compressed(os.path.join(HERE, "wedge_2d.npz"), **wedge_2d())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
