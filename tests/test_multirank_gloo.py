"""world_size-2 `gloo` tests (CPU) of the N>1 host logic.

Each rank owns the boxes fvs_distribute_boxes assigns to it, executes the halo
plan of the C ABI (fvs_halo_plan) -- local copies in place, remote copies as
point-to-point messages in plan order, exactly how the CUDA context pairs its
ncclSend/ncclRecv -- and reduces its stable dt with a MIN all-reduce (the
collective fvs_min_all issues).  The result must equal the single-process
oracle: FillBoundary of the whole level and the level-wide stable dt.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RANK, G = 3, 4
DOMAIN = ([0, 0, 0], [15, 11, 7])
BOXES = [([0, 0, 0], [7, 5, 7]), ([8, 0, 0], [15, 5, 7]),
         ([0, 6, 0], [7, 11, 7]), ([8, 6, 0], [15, 11, 7])]
PERIODIC = (1, 0, 0)
DX = [1.0 / 16, 1.0 / 12, 1.0 / 8]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _region(c, box_lo, shift):
    lo, hi = c["lo"], c["hi"]
    return tuple([slice(None)] + [slice(lo[d] - shift[d] - box_lo[d], hi[d] - shift[d] - box_lo[d] + 1)
                                  for d in (2, 1, 0)])


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from finitevolumesolver_b200 import capi
    from helpers import smooth_state
    from oracle import oracle as orc

    owner = capi.distribute_boxes(RANK, BOXES, world)
    plan = capi.halo_plan(RANK, DOMAIN[0], DOMAIN[1], PERIODIC, BOXES, owner, G)
    fab_lo = {b: [BOXES[b][0][d] - G for d in range(3)] for b in range(len(BOXES))}
    fab_hi = {b: [BOXES[b][1][d] + G for d in range(3)] for b in range(len(BOXES))}
    mine = {b: smooth_state(RANK, fab_lo[b], fab_hi[b], DX, seed=b)
            for b in range(len(BOXES)) if owner[b] == rank}
    before = {b: a.copy() for b, a in mine.items()}

    # the exchange: sources are read from the state BEFORE the fill
    zero3 = [0, 0, 0]
    reqs, recvs = [], []
    for i, c in enumerate(plan):
        so, do = c["src_owner"], c["dst_owner"]
        if so == rank and do == rank:
            mine[c["dst_box"]][_region(c, fab_lo[c["dst_box"]], zero3)] = \
                before[c["src_box"]][_region(c, fab_lo[c["src_box"]], c["shift"])]
        elif so == rank:
            t = torch.from_numpy(np.ascontiguousarray(
                before[c["src_box"]][_region(c, fab_lo[c["src_box"]], c["shift"])]))
            reqs.append(dist.isend(t, do, tag=i))
        elif do == rank:
            shape = [RANK + 4] + [c["hi"][d] - c["lo"][d] + 1 for d in (2, 1, 0)]
            t = torch.empty(shape, dtype=torch.float64)
            reqs.append(dist.irecv(t, so, tag=i))
            recvs.append((c, t))
    for r in reqs:
        r.wait()
    for c, t in recvs:
        mine[c["dst_box"]][_region(c, fab_lo[c["dst_box"]], zero3)] = t.numpy()

    # single-process oracle of the whole level
    lvl = orc.Level(RANK, DOMAIN[0], DOMAIN[1], DX, BOXES, scratch_gcw=G, flux_gcw=G - 2,
                    periodic=PERIODIC)
    for b in range(len(BOXES)):
        arr, lo, hi = lvl.scratch(b)
        assert lo == fab_lo[b] and hi == fab_hi[b]
        arr[...] = smooth_state(RANK, lo, hi, DX, seed=b)
    lvl.fill_boundary()
    ok = all(np.array_equal(mine[b], lvl.scratch(b)[0]) for b in mine)

    # stable dt: rank-local min over owned boxes, then MIN all-reduce
    level_dt = min(lvl.compute_stable_dt(d) for d in range(RANK))
    local = np.inf
    for b in mine:
        one = orc.Level(RANK, DOMAIN[0], DOMAIN[1], DX, [BOXES[b]], scratch_gcw=G, flux_gcw=G - 2,
                        periodic=PERIODIC)
        one.scratch(0)[0][...] = mine[b]
        local = min(local, min(one.compute_stable_dt(d) for d in range(RANK)))
    t = torch.tensor([local], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dt_ok = float(t.item()) == level_dt

    # every remote copy is seen by both of its ranks, in the same order
    remote = [(i, c["src_owner"], c["dst_owner"]) for i, c in enumerate(plan)
              if c["src_owner"] != c["dst_owner"]]
    gathered = [None] * world
    dist.all_gather_object(gathered, remote)
    same_plan = all(g == gathered[0] for g in gathered)

    with open(os.path.join(tmpdir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{int(ok)} {int(dt_ok)} {int(same_plan)} {len(mine)} {len(remote)}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_halo_exchange_and_dt_min_match_single_process_oracle(tmp_path):
    from oracle import oracle as orc
    orc.build()
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, dt_ok, same_plan, nmine, nremote = map(int, (tmp_path / f"rank{r}.txt").read_text().split())
        assert nmine == 2 and nremote > 0
        assert ok, f"rank {r}: ghost cells differ from the oracle's FillBoundary"
        assert dt_ok, f"rank {r}: reduced stable dt differs from the level dt"
        assert same_plan


def test_bench_box_layout_is_one_box_per_rank():
    """bench.py --gpus N: N boxes, box b on rank b; weak scaling grows the domain with N,
    strong scaling cuts the same domain into N boxes; the parity sub-box of bench.py lies
    inside the domain and touches every rank boundary it can."""
    import importlib
    bench = importlib.import_module("bench")
    for n in (1, 2, 4, 8):
        domain_n, boxes, parts, box_n = bench.grid_of_boxes(n, 16)
        assert len(boxes) == n and parts[0] * parts[1] * parts[2] == n
        assert domain_n[0] * domain_n[1] * domain_n[2] == n * 16 ** 3
        cells = set()
        for lo, hi in boxes:
            assert all(hi[d] - lo[d] + 1 == 16 for d in range(3))
            cells.add(tuple(lo))
        assert len(cells) == n
        sdom, sboxes, sparts, sbox_n = bench.grid_of_boxes(n, 16, strong=True)
        assert sdom == [16, 16, 16] and len(sboxes) == n
        assert sum((hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1)
                   for lo, hi in sboxes) == 16 ** 3
