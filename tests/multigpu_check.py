"""Multi-GPU parity check (run under torchrun, one rank per GPU; not a pytest
file because it needs N real GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Every rank owns one box of a 3-D level, exchanges wide halos with ncclSend /
ncclRecv and reduces the stable dt with ncclAllReduce; the result of each
rank's box and the dt sequence are compared bit for bit with the CPU oracle
run on the whole level."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import assert_bitwise_equal, riemann_state  # noqa: E402


def main():
    from finitevolumesolver_b200 import capi
    from oracle import oracle as orc
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 24
    parts = [1, 1, 1]
    d = 0
    w = world
    while w > 1:
        parts[d % 3] *= 2
        w //= 2
        d += 1
    boxes = []
    for k in range(parts[2]):
        for j in range(parts[1]):
            for i in range(parts[0]):
                lo = [i * n, j * n, k * n]
                boxes.append((lo, [lo[0] + n - 1, lo[1] + n - 1, lo[2] + n - 1]))
    dom = [parts[q] * n for q in range(3)]
    dx = [2.0 / dom[0]] * 3
    origin = [-1.0, -1.0 * dom[1] / dom[0], -1.0 * dom[2] / dom[0]]
    periodic = (1, 0, 0)
    ctx = capi.Context(3, [0] * 3, [q - 1 for q in dom], dx, boxes, scratch_gcw=8, flux_gcw=6,
                       periodic=periodic, device=local, comm_rank=rank, comm_size=world,
                       owner=list(range(world)))
    idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
    cpu = orc.Level(3, [0] * 3, [q - 1 for q in dom], dx, boxes, scratch_gcw=8, flux_gcw=6,
                    periodic=periodic)
    cpu.set_method(0, 1, 2)
    for q in (1, 2):
        for s in (0, 1):
            ctx.add_bc(q, s, capi.BC_REFLECTIVE if q == 1 else capi.BC_TRANSMISSIVE)
            cpu.add_bc(q, s, orc.BC_REFLECTIVE if q == 1 else orc.BC_TRANSMISSIVE)
    for b in range(len(boxes)):
        arr, lo, hi = cpu.scratch(b)
        arr[...] = riemann_state(3, lo, hi, dx, origin, radius=0.45)
        if b == rank:
            ctx.upload_state(b, arr)
    cpu.fill_ghost_layer()
    ctx.fill_ghost_layer()
    assert_bitwise_equal(ctx.download_state(rank), cpu.scratch(rank)[0], f"rank {rank}: halo")
    for step in range(6):
        dt_c = min(cpu.compute_stable_dt(q) for q in range(3))
        dt_g = ctx.compute_stable_dt_all()
        assert dt_c == dt_g, (rank, step, dt_c, dt_g)
        ok_c, _, dts_c = cpu.advance_hierarchy(0.5 * dt_c, orc.SPLIT_STRANG)
        ok_g, _, dts_g = ctx.advance_hierarchy(0.5 * dt_g, capi.SPLIT_STRANG)
        assert ok_c and ok_g and dts_c == dts_g, (rank, step, dts_c, dts_g)
    assert_bitwise_equal(ctx.download_state(rank), cpu.scratch(rank)[0], f"rank {rank}: 6 steps")
    v = ctx.min_all(float(rank + 3))
    assert v == 3.0, v
    ctx.close()
    dist.barrier()
    eb_check(capi, orc, rank, world, local)
    dist.barrier()
    amr_check(capi, orc, rank, world, local)
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: {world} ranks, halo + 6 Strang steps bit-identical to the "
              f"oracle, identical dt sequence, min_all ok; cut-cell level over {world} ranks "
              f"bit-identical too; 3-level subcycled hierarchy with coarse and fine boxes on "
              f"different ranks bit-identical")
    dist.destroy_process_group()


def amr_check(capi, orc, rank, world, local):
    """Three nested levels whose boxes are dealt round robin over the ranks, so
    that average_down, pc_interp, CrseInit and Reflux all cross rank borders;
    fine-first subcycling against the oracle's hierarchy."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import complete_from_prim
    n0 = [32, 24]
    level_boxes = [
        [([0, 0, 0], [15, 23, 0]), ([16, 0, 0], [31, 23, 0])],
        [([16, 8, 0], [31, 39, 0]), ([32, 8, 0], [47, 23, 0]), ([32, 24, 0], [47, 39, 0])],
        [([48, 40, 0], [71, 63, 0]), ([72, 32, 0], [87, 55, 0])],
    ]
    owners = [[b % world for b in range(2)], [(b + 1) % world for b in range(3)],
              [b % world for b in range(2)]]
    cpus, gpus = [], []
    ctx = None
    for l, boxes in enumerate(level_boxes):
        f = 2 ** l
        hi = [n0[0] * f - 1, n0[1] * f - 1, 0]
        dx = [2.0 / (n0[0] * f), 2.0 / (n0[1] * f), 1.0]
        cpu = orc.Level(2, [0] * 3, hi, dx, boxes, scratch_gcw=8, flux_gcw=6)
        if l == 0:
            ctx = capi.Context(2, [0] * 3, hi, dx, boxes, scratch_gcw=8, flux_gcw=6, device=local,
                               comm_rank=rank, comm_size=world, owner=owners[0])
            idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
            gpu = ctx
        else:
            gpu = ctx.define_level(l, [0] * 3, hi, dx, boxes, scratch_gcw=8, flux_gcw=6,
                                   owner=owners[l])
        for lv in (cpu, gpu):
            for d in range(2):
                for s in (0, 1):
                    lv.add_bc(d, s, 1)
        for b in range(len(boxes)):
            arr, lo, hi_b = cpu.scratch(b)
            x = -1.0 + (np.arange(lo[0], hi_b[0] + 1) + 0.5) * dx[0]
            y = -1.0 + (np.arange(lo[1], hi_b[1] + 1) + 0.5) * dx[1]
            upper = (x[None, None, :] + y[None, :, None]) > 0.0
            rho = np.where(upper, 0.125, 1.0)
            p = np.where(upper, 10132.5, 101325.0)
            arr[...] = complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, p)
            if owners[l][b] == rank:
                gpu.upload_state(b, arr)
        cpus.append(cpu)
        gpus.append(gpu)
    hier = orc.Hierarchy(cpus)
    cpus[0].fill_ghost_layer()
    gpus[0].fill_ghost_layer()
    for l in range(1, 3):
        orc.CoarseFine(cpus[l], cpus[l - 1]).fill_ghost_layer_two_levels()
        ctx.fill_ghost_layer_two_levels(l, l - 1)
    for step in range(3):
        dt_c = hier.compute_stable_dt()
        dt_g = ctx.compute_stable_dt_hierarchy()
        assert dt_c == dt_g, (rank, step, dt_c, dt_g)
        ok_c, _, dts_c = hier.advance_hierarchy(0.5 * dt_c, orc.SPLIT_STRANG)
        ok_g, _, dts_g = ctx.advance_hierarchy(0.5 * dt_g, capi.SPLIT_STRANG)
        assert ok_c and ok_g and dts_c == dts_g, (rank, step, dts_c, dts_g)
        for l in range(3):
            for b in range(len(level_boxes[l])):
                if owners[l][b] == rank:
                    assert_bitwise_equal(gpus[l].download_state(b), cpus[l].scratch(b)[0],
                                         f"rank {rank} step {step} level {l} box {b}")
    ctx.close()


def eb_check(capi, orc, rank, world, local):
    """A 2-D cut-cell level (wedge) whose boxes are spread over the ranks:
    KbnCutCellMethod + NCCL halo + dt all-reduce, against the oracle."""
    from helpers import complete_from_prim, eb_fab_type, eb_geometry, polygon_phi
    nbx = world
    n = (24 * nbx, 32)
    dx = [0.15 / n[0], 0.10 / n[1], 1.0]
    boxes = [([24 * b, 0, 0], [24 * b + 23, n[1] - 1, 0]) for b in range(nbx)]
    ctx = capi.Context(2, [0] * 3, [n[0] - 1, n[1] - 1, 0], dx, boxes, scratch_gcw=8, flux_gcw=6,
                       device=local, comm_rank=rank, comm_size=world, owner=list(range(world)))
    idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
    cpu = orc.Level(2, [0] * 3, [n[0] - 1, n[1] - 1, 0], dx, boxes, scratch_gcw=8, flux_gcw=6)
    cpu.set_method(0, 1, 2)
    cpu.enable_eb()
    ctx.enable_eb()
    phi = polygon_phi([(0.04, 0.05), (0.10, 0.075), (0.10, 0.025)])
    for d in range(2):
        for s in (0, 1):
            ctx.add_bc(d, s, capi.BC_TRANSMISSIVE)
            cpu.add_bc(d, s, orc.BC_TRANSMISSIVE)
    for b in range(nbx):
        arr, lo, hi = cpu.scratch(b)
        g = eb_geometry(2, lo, hi, dx, [0.0] * 3, phi)
        cpu.set_eb(b, g, eb_fab_type(g))
        x = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
        rho = np.ones(arr.shape[1:])
        p = np.where(x < 0.02, 5 * 101325.0, 101325.0)[None, None, :] + np.zeros_like(rho)
        arr[...] = complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, p) * (g["alpha"][0] > 0)
        if b == rank:
            ctx.set_eb(b, g, eb_fab_type(g))
            ctx.upload_state(b, arr)
    cpu.fill_ghost_layer()
    ctx.fill_ghost_layer()
    for step in range(8):
        dt_c = min(cpu.compute_stable_dt(q) for q in range(2))
        dt_g = ctx.compute_stable_dt_all()
        assert dt_c == dt_g, (rank, step, dt_c, dt_g)
        ok_c, _, dts_c = cpu.advance_hierarchy(0.4 * dt_c, orc.SPLIT_STRANG)
        ok_g, _, dts_g = ctx.advance_hierarchy(0.4 * dt_g, capi.SPLIT_STRANG)
        assert ok_c and ok_g and dts_c == dts_g, (rank, step, dts_c, dts_g)
    assert_bitwise_equal(ctx.download_state(rank), cpu.scratch(rank)[0], f"rank {rank}: EB level")
    ctx.close()


if __name__ == "__main__":
    main()
