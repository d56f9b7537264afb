"""CPU tests of the product's host side: the C-ABI library loads and exports
every symbol include/fvs_b200.h declares, refuses to run without a GPU, and
its box arithmetic (stable-dt face set, halo plan, box distribution) agrees
with the oracle's restatement of the AMReX algorithms."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fvs():
    from finitevolumesolver_b200 import build, capi
    build.build()
    capi.lib()
    return capi


def test_library_exports_every_declared_symbol(fvs):
    header = open(os.path.join(ROOT, "include", "fvs_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fvs_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(fvs.SYMBOLS)
    L = fvs.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.fvs_version() == 1


def test_product_library_does_not_contain_the_oracle(fvs):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", fvs.LIB_PATH], capture_output=True,
                         text=True).stdout
    assert "fvso_" not in out
    assert "fvs_sweep" in out


def test_create_fails_loudly_without_gpu(fvs):
    if fvs.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fvs.FvsError, match="no CUDA device"):
        fvs.Context(1, [0], [15], [1.0], [([0], [15])], scratch_gcw=2, flux_gcw=0)


@pytest.mark.parametrize("length", [8, 16, 17, 23, 64, 100, 7, 3])
@pytest.mark.parametrize("w", [1, 2])
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_stable_dt_face_mask_matches_oracle_tiles(fvs, oracle, length, w, direction):
    """FluxMethodAdapter.hpp:82-98: tiles of 1024000 x 8 x 8, growntilebox,
    stencil positions inside each tile."""
    rank, g = 3, 4
    lo = [0, 0, 0]
    hi = [11, 11, 11]
    lo[direction], hi[direction] = 5, 5 + length - 1
    mask = fvs.stable_dt_face_mask(rank, lo, hi, g, w, direction)
    assert len(mask) == length + 2 * g
    expect = np.zeros(length + 2 * g, dtype=np.uint8)
    for tile_lo, tile_hi, grown_lo, grown_hi in oracle.tiles(rank, lo, hi, [1024000, 8, 8],
                                                             [g, g, g]):
        a, b = grown_lo[direction], grown_hi[direction]
        for s in range(a, b - 2 * w + 2):
            expect[s + w - 1 - (lo[direction] - g)] = 1
    assert np.array_equal(mask, expect)


def _brute_fill_boundary(rank, domain_lo, domain_hi, periodic, boxes, g, arrays):
    """Ghost cells inside another box's valid region (or a periodic image)."""
    out = [a.copy() for a in arrays]
    dom = [domain_hi[d] - domain_lo[d] + 1 if d < rank else 1 for d in range(3)]
    for i, (lo_i, hi_i) in enumerate(boxes):
        flo = [lo_i[d] - g if d < rank else 0 for d in range(3)]
        fhi = [hi_i[d] + g if d < rank else 0 for d in range(3)]
        for k in range(flo[2], fhi[2] + 1):
            for j in range(flo[1], fhi[1] + 1):
                for ii in range(flo[0], fhi[0] + 1):
                    idx = [ii, j, k]
                    if all(lo_i[d] <= idx[d] <= hi_i[d] for d in range(3)):
                        continue
                    for jb, (lo_j, hi_j) in enumerate(boxes):
                        hit = None
                        rng = [(-1, 0, 1) if (d < rank and periodic[d]) else (0,) for d in range(3)]
                        for sz in rng[2]:
                            for sy in rng[1]:
                                for sx in rng[0]:
                                    s = [sx * dom[0], sy * dom[1], sz * dom[2]]
                                    src = [idx[d] - s[d] for d in range(3)]
                                    if all(lo_j[d] <= src[d] <= hi_j[d] for d in range(3)):
                                        hit = src
                        if hit is not None:
                            slo = [lo_j[d] - g if d < rank else 0 for d in range(3)]
                            out[i][:, k - flo[2], j - flo[1], ii - flo[0]] = \
                                arrays[jb][:, hit[2] - slo[2], hit[1] - slo[1], hit[0] - slo[0]]
    return out


def _apply_plan(plan, rank, boxes, g, arrays):
    out = [a.copy() for a in arrays]
    for c in plan:
        dlo = [boxes[c["dst_box"]][0][d] - g if d < rank else 0 for d in range(3)]
        slo = [boxes[c["src_box"]][0][d] - g if d < rank else 0 for d in range(3)]
        lo, hi, s = c["lo"], c["hi"], c["shift"]
        dst = out[c["dst_box"]]
        src = arrays[c["src_box"]]
        dst[:, lo[2] - dlo[2]:hi[2] - dlo[2] + 1, lo[1] - dlo[1]:hi[1] - dlo[1] + 1,
            lo[0] - dlo[0]:hi[0] - dlo[0] + 1] = \
            src[:, lo[2] - s[2] - slo[2]:hi[2] - s[2] - slo[2] + 1,
                lo[1] - s[1] - slo[1]:hi[1] - s[1] - slo[1] + 1,
                lo[0] - s[0] - slo[0]:hi[0] - s[0] - slo[0] + 1]
    return out


@pytest.mark.parametrize("periodic", [(0, 0, 0), (1, 0, 0), (1, 1, 0)])
def test_halo_plan_reproduces_fill_boundary(fvs, periodic):
    rank, g = 2, 3
    boxes = [([0, 0, 0], [7, 5, 0]), ([8, 0, 0], [15, 5, 0]), ([0, 6, 0], [15, 11, 0])]
    rng = np.random.default_rng(1)
    arrays = [rng.random((2, 1, hi[1] - lo[1] + 1 + 2 * g, hi[0] - lo[0] + 1 + 2 * g))
              for lo, hi in boxes]
    plan = fvs.halo_plan(rank, [0, 0], [15, 11], periodic, boxes, [0, 1, 0], g)
    got = _apply_plan(plan, rank, boxes, g, arrays)
    ref = _brute_fill_boundary(rank, [0, 0, 0], [15, 11, 0], periodic, boxes, g, arrays)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    assert {(c["src_owner"], c["dst_owner"]) for c in plan} >= {(0, 1), (1, 0), (0, 0)}


def test_distribute_boxes_balances_and_keeps_neighbours(fvs):
    boxes = []
    for k in range(2):
        for j in range(2):
            for i in range(2):
                boxes.append(([i * 8, j * 8, k * 8], [i * 8 + 7, j * 8 + 7, k * 8 + 7]))
    for nranks in (1, 2, 4, 8):
        owner = fvs.distribute_boxes(3, boxes, nranks)
        counts = np.bincount(owner, minlength=nranks)
        assert counts.min() == counts.max() == 8 // nranks
    owner = fvs.distribute_boxes(3, boxes, 3)
    assert sorted(set(owner)) == [0, 1, 2]
