"""Tag -> regrid -> step on the GPU against the oracle (BASELINE config 3, SURVEY W3 in
small): the 2-D diagonal Sod problem of examples/AMReX/SodShockTube2d.cpp:57-64 on three
levels, GradientDetector(density, 1e-2) + TagBuffer (:113-114,126), regridding before every
coarse step (IntegratorContext::PreAdvanceLevel, src/AMReX/IntegratorContext.cpp:601-627 ->
GriddingAlgorithm::RegridAllFinerlevels, src/AMReX/GriddingAlgorithm.cpp:86-110), fine-first
subcycling with flux correction.  The tags come from the device kernels, the boxes from
fvs_cluster_tags (the stand-in for AmrCore::regrid); the oracle runs on the same boxes.
Bit-identical states on every level, identical dt, identical tags."""
import numpy as np
import pytest

from helpers import assert_bitwise_equal, complete_from_prim

pytestmark = pytest.mark.gpu

RANK = 2
N0 = (64, 64)
GCW = (8, 6)
METHOD = (0, 1, 2)
BF, MAXGRID, NEST, BUFFER = 8, 32, 1, 2


@pytest.fixture(scope="module")
def fvs():
    from finitevolumesolver_b200 import capi
    assert capi.device_count() > 0, "no CUDA device"
    return capi


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def level_geometry(l):
    f = 2 ** l
    hi = [N0[0] * f - 1, N0[1] * f - 1, 0]
    dx = [2.0 / (N0[0] * f), 2.0 / (N0[1] * f), 1.0]
    return hi, dx


def new_cpu_level(orc, l, boxes):
    hi, dx = level_geometry(l)
    L = orc.Level(RANK, [0] * 3, hi, dx, boxes, scratch_gcw=GCW[0], flux_gcw=GCW[1])
    L.set_method(*METHOD)
    for d in range(RANK):
        for s in (0, 1):
            L.add_bc(d, s, 1)
    return L


def sod(lo, hi, dx):
    x = -1.0 + (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    y = -1.0 + (np.arange(lo[1], hi[1] + 1) + 0.5) * dx[1]
    upper = (x[None, None, :] + y[None, :, None]) > 0.0
    rho = np.where(upper, 0.125, 1.0)
    p = np.where(upper, 10132.5, 101325.0)
    return complete_from_prim(RANK, rho, [np.zeros_like(rho)] * 2, p)


def level_tags(fvs, orc, l, cpu, gpu):
    """Dense tag map of level l from the device kernels (checked against the oracle's)."""
    hi, _ = level_geometry(l)
    dense = np.zeros((1, hi[1] + 1, hi[0] + 1), dtype=np.int8)
    for b, (lo, h) in enumerate(cpu.boxes):
        t_g = gpu.tag_gradient(b, [(0, 1e-2)])
        t_c = cpu.tag_gradient(b, [(0, 1e-2)])
        assert np.array_equal(t_g, t_c), f"tags of level {l} box {b}"
        dense[0, lo[1]:h[1] + 1, lo[0]:h[0] + 1] |= t_g.reshape(1, h[1] - lo[1] + 1, h[0] - lo[0] + 1)[0]
    buffered = dense.copy()
    gpu.tag_buffer(buffered, RANK, BUFFER)
    ref = dense.copy()
    orc.tag_buffer(ref, RANK, BUFFER)
    assert np.array_equal(buffered, ref)
    # the level-wide entry point gives the same map with one synchronisation
    whole = gpu.tag_level([hi[0] + 1, hi[1] + 1, 1], [(0, 1e-2)], buffer_width=BUFFER)
    assert np.array_equal(whole, ref)
    return buffered


def regrid(fvs, orc, ctx, cpus, gpus, max_levels, init=False):
    """Levels 1 .. max_levels-1 from the tags of the level below, coarse to fine."""
    changed = 0
    l = 0
    while l + 1 < max_levels and l < len(cpus):
        tags = level_tags(fvs, orc, l, cpus[l], gpus[l])
        hi, _ = level_geometry(l)
        boxes = fvs.cluster_tags(RANK, [0] * 3, hi, [0] * 3, cpus[l].boxes, tags[0][None],
                                 blocking_factor=BF, max_grid_size=MAXGRID, nesting_buffer=NEST)
        if not boxes:
            ctx.regrid_level(l + 1, [])
            del cpus[l + 1:], gpus[l + 1:]
            break
        old_cpu = cpus[l + 1] if l + 1 < len(cpus) else None
        if old_cpu is None or [list(map(list, b)) for b in old_cpu.boxes] != [list(map(list, b)) for b in boxes]:
            changed += 1
        fresh = new_cpu_level(orc, l + 1, boxes)
        orc.regrid_fill(fresh, old_cpu, cpus[l])
        g = ctx.regrid_level(l + 1, boxes)
        if init:  # MakeNewLevelFromScratch: the initial data on the new level
            fhi, fdx = level_geometry(l + 1)
            for b in range(len(boxes)):
                arr, lo, h = fresh.scratch(b)
                arr[...] = sod(lo, h, fdx)
                g.upload_state(b, arr)
            fresh.fill_ghost_layer()
            g.fill_ghost_layer()
        if l + 1 < len(cpus):
            cpus[l + 1], gpus[l + 1] = fresh, g
        else:
            cpus.append(fresh)
            gpus.append(g)
        l += 1
    return changed


def compare(cpus, gpus, what):
    for l, (c, g) in enumerate(zip(cpus, gpus)):
        for b in range(len(c.boxes)):
            assert_bitwise_equal(g.download_state(b), c.scratch(b)[0], f"{what} level {l} box {b}")


@pytest.mark.parametrize("steps", [int(__import__("os").environ.get("FVS_REGRID_STEPS", "100"))])
def test_tag_regrid_step_bitwise(oracle, fvs, steps):
    orc = oracle
    hi0, dx0 = level_geometry(0)
    boxes0 = [([0, 0, 0], [31, 31, 0]), ([32, 0, 0], [63, 31, 0]), ([0, 32, 0], [31, 63, 0]),
              ([32, 32, 0], [63, 63, 0])]
    c0 = new_cpu_level(orc, 0, boxes0)
    ctx = fvs.Context(RANK, [0] * 3, hi0, dx0, boxes0, scratch_gcw=GCW[0], flux_gcw=GCW[1],
                      base=METHOD[0], recon=METHOD[1], limiter=METHOD[2])
    for d in range(RANK):
        for s in (0, 1):
            ctx.add_bc(d, s, 1)
    for b in range(len(boxes0)):
        arr, lo, h = c0.scratch(b)
        arr[...] = sod(lo, h, dx0)
        ctx.upload_state(b, arr)
    c0.fill_ghost_layer()
    ctx.fill_ghost_layer()
    cpus, gpus = [c0], [ctx]
    regrid(fvs, orc, ctx, cpus, gpus, 3, init=True)
    assert len(cpus) == 3, "the diagonal discontinuity is refined twice"
    compare(cpus, gpus, "initial hierarchy")
    regrids = 0
    for step in range(steps):
        regrids += regrid(fvs, orc, ctx, cpus, gpus, 3)
        compare(cpus, gpus, f"after regrid, step {step}")
        hier = orc.Hierarchy(cpus)
        dt_c = hier.compute_stable_dt()
        dt_g = ctx.compute_stable_dt_hierarchy()
        assert dt_c == dt_g, (step, dt_c, dt_g)
        ok_c, _, dts_c = hier.advance_hierarchy(0.5 * dt_c, 1)
        ok_g, _, dts_g = ctx.advance_hierarchy(0.5 * dt_g, 1)
        assert ok_c and ok_g and dts_c == dts_g, (step, dts_c, dts_g)
        compare(cpus, gpus, f"step {step}")
    # the waves moved: the boxes of the finer levels changed on the way
    assert regrids >= 3, regrids
    assert len(cpus) == 3
    ctx.close()
