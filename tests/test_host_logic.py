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


def _cells(boxes):
    s = set()
    for lo, hi in boxes:
        for k in range(lo[2], hi[2] + 1):
            for j in range(lo[1], hi[1] + 1):
                for i in range(lo[0], hi[0] + 1):
                    s.add((i, j, k))
    return s


@pytest.mark.parametrize("rank", [2, 3])
def test_cluster_tags_contract(fvs, rank):
    """The regrid hand-off (stand-in for AmrCore::regrid behind
    src/AMReX/GriddingAlgorithm.cpp:86-110): boxes are disjoint, aligned to the blocking
    factor, at most max_grid_size long, properly nested, and cover every tagged cell that
    can be refined."""
    rng = np.random.default_rng(5)
    n = [32, 24, 16 if rank == 3 else 1]
    hi = [n[0] - 1, n[1] - 1, n[2] - 1]
    # level boxes: an L-shaped part of the domain
    boxes = [([0, 0, 0], [15, hi[1], hi[2]]), ([16, 8, 0], [31, hi[1], hi[2]])]
    tags = (rng.random((n[2], n[1], n[0])) < 0.02).astype(np.int8)
    bf, mg, nest = 8, 16, 1
    out = fvs.cluster_tags(rank, [0, 0, 0], hi, [0, 0, 0], boxes, tags, blocking_factor=bf,
                           max_grid_size=mg, nesting_buffer=nest)
    assert len(out) > 1
    total = 0
    for lo, h in out:
        for d in range(rank):
            assert lo[d] % bf == 0 and (h[d] + 1) % bf == 0
            assert h[d] - lo[d] + 1 <= mg
        for d in range(rank, 3):
            assert lo[d] == 0 and h[d] == 0
        total += np.prod([h[d] - lo[d] + 1 for d in range(3)])
    fine_cells = _cells(out)
    assert len(fine_cells) == total, "boxes overlap"
    level_cells = _cells(boxes)
    r = [2 if d < rank else 1 for d in range(3)]
    refined_parents = {(i // r[0], j // r[1], k // r[2]) for (i, j, k) in fine_cells}
    rng_nb = [range(-nest, nest + 1) if d < rank else range(1) for d in range(3)]
    for (i, j, k) in refined_parents:  # proper nesting
        for dk in rng_nb[2]:
            for dj in rng_nb[1]:
                for di in rng_nb[0]:
                    q = (i + di, j + dj, k + dk)
                    inside_domain = all(0 <= q[d] <= hi[d] for d in range(3))
                    assert (not inside_domain) or q in level_cells
    bc = bf // 2
    for k, j, i in zip(*np.nonzero(tags)):
        blk = [(i // bc) * bc, (j // bc) * bc, (k // bc) * bc if rank == 3 else 0]
        nestable = True
        for kk in range(blk[2] - (nest if rank == 3 else 0), blk[2] + (bc if rank == 3 else 1) + (nest if rank == 3 else 0)):
            for jj in range(blk[1] - nest, blk[1] + bc + nest):
                for ii in range(blk[0] - nest, blk[0] + bc + nest):
                    q = (ii, jj, kk)
                    if all(0 <= q[d] <= hi[d] for d in range(3)) and q not in level_cells:
                        nestable = False
        assert ((i, j, k) in refined_parents) == nestable, (i, j, k)
    # deterministic
    assert out == fvs.cluster_tags(rank, [0, 0, 0], hi, [0, 0, 0], boxes, tags, blocking_factor=bf,
                                   max_grid_size=mg, nesting_buffer=nest)
    # no tags, no boxes
    assert fvs.cluster_tags(rank, [0, 0, 0], hi, [0, 0, 0], boxes, 0 * tags) == []


def test_cutcell_load_balance_weights(fvs):
    """cutcell::GriddingAlgorithm::LoadBalance (src/AMReX/cutcell/GriddingAlgorithm.cpp:190-211):
    regular cells weigh 1, cells of single-valued fabs 6, covered cells 0.001; knapsack."""
    boxes = [([8 * i, 0, 0], [8 * i + 7, 7, 0]) for i in range(8)]
    fab = [0, 1, 0, -1, 0, 1, 0, 0]          # weights 64, 6*54+0.01=324.01, 64, 0.064, 64, 384, 64, 64
    cov = [0, 10, 0, 64, 0, 0, 0, 0]
    owner = fvs.distribute_boxes_cutcell(2, boxes, fab, cov, 2)
    w = [64, 6 * 54 + 0.001 * 10, 64, 0.001 * 64, 64, 6 * 64, 64, 64]
    load = [sum(w[i] for i in range(8) if owner[i] == r) for r in range(2)]
    assert owner[1] != owner[5], "the two cut fabs go to different ranks"
    assert abs(load[0] - load[1]) <= 64.0
    # equal-count assignment would put boxes 0-3 on rank 0: 388 vs 640 -- the weights matter
    assert max(load) < 560
    # a single rank takes everything
    assert fvs.distribute_boxes_cutcell(2, boxes, fab, cov, 1) == [0] * 8
