"""Parity of the coarse-fine kernels (two levels on one GPU) against the CPU
oracle: pc_interp ghost fill, flux registers (FineAdd, CrseInit, Reflux,
cleared internal borders), average_down.  Bit-exact; the arithmetic restates
AMReX's published algorithms ("parity unpinned" against a real AMReX build)."""
import numpy as np
import pytest

from helpers import assert_bitwise_equal, smooth_state

pytestmark = pytest.mark.gpu


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


def two_levels(orc, fvs, rank, periodic=False):
    """periodic: both levels periodic in every direction and fine boxes ON the
    periodic boundaries, two of them periodic neighbours of each other (ghost
    parents, CrseInit faces and Reflux cells come from the coarse level's
    periodic images; amrex copies with geom.periodicity())."""
    if rank == 2:
        nc = [32, 24, 1]
        cboxes = [([0, 0, 0], [15, 23, 0]), ([16, 0, 0], [31, 23, 0])]
        fboxes = [([16, 8, 0], [31, 31, 0]), ([32, 8, 0], [47, 23, 0]), ([32, 24, 0], [47, 39, 0])]
        if periodic:
            fboxes = [([0, 8, 0], [15, 31, 0]), ([48, 8, 0], [63, 23, 0]), ([24, 0, 0], [39, 15, 0])]
    else:
        nc = [16, 12, 12]
        cboxes = [([0, 0, 0], [15, 11, 11])]
        fboxes = [([8, 4, 4], [19, 15, 15]), ([20, 4, 4], [27, 15, 19])]
        if periodic:
            fboxes = [([0, 4, 4], [11, 15, 15]), ([20, 4, 16], [31, 15, 23])]
    dxc = [1.0 / nc[d] if d < rank else 1.0 for d in range(3)]
    dxf = [0.5 * dxc[d] if d < rank else 1.0 for d in range(3)]
    chi = [nc[d] - 1 if d < rank else 0 for d in range(3)]
    fhi = [2 * nc[d] - 1 if d < rank else 0 for d in range(3)]
    gcw = (4, 2)
    kw = dict(scratch_gcw=gcw[0], flux_gcw=gcw[1])
    if periodic:
        kw["periodic"] = [1 if d < rank else 0 for d in range(3)]
    ccpu = orc.Level(rank, [0] * 3, chi, dxc, cboxes, **kw)
    fcpu = orc.Level(rank, [0] * 3, fhi, dxf, fboxes, **kw)
    cgpu = fvs.Context(rank, [0] * 3, chi, dxc, cboxes, **kw)
    fgpu = cgpu.define_level(1, [0] * 3, fhi, dxf, fboxes, **kw)
    for lv in (ccpu, fcpu, cgpu, fgpu):
        for d in range(rank):
            for s in (0, 1):
                if not periodic:
                    lv.add_bc(d, s, 1 if d == 0 else 0)
    for cpu, gpu, dx, seed in ((ccpu, cgpu, dxc, 1), (fcpu, fgpu, dxf, 2)):
        for b in range(len(cpu.boxes)):
            arr, lo, hi = cpu.scratch(b)
            arr[...] = smooth_state(rank, lo, hi, dx, seed=seed)
            gpu.upload_state(b, arr)
    cf = orc.CoarseFine(fcpu, ccpu)
    return ccpu, fcpu, cgpu, fgpu, cf


def compare(cpu, gpu, what):
    for b in range(len(cpu.boxes)):
        assert_bitwise_equal(gpu.download_state(b), cpu.scratch(b)[0], f"{what} box {b}")


@pytest.mark.parametrize("rank,periodic", [(2, False), (3, False), (2, True), (3, True)])
def test_fill_ghost_layer_two_levels(oracle, fvs, rank, periodic):
    ccpu, fcpu, cgpu, fgpu, cf = two_levels(oracle, fvs, rank, periodic)
    ccpu.fill_ghost_layer()
    cgpu.fill_ghost_layer()
    cf.fill_ghost_layer_two_levels()
    cgpu.fill_ghost_layer_two_levels(1, 0)
    compare(fcpu, fgpu, "pc_interp ghost fill")
    cgpu.close()


@pytest.mark.parametrize("rank", [2, 3])
def test_coarsen_conservatively(oracle, fvs, rank):
    ccpu, fcpu, cgpu, fgpu, cf = two_levels(oracle, fvs, rank)
    before = cgpu.download_state(0).copy()
    cf.coarsen_conservatively()
    cgpu.coarsen_conservatively(1, 0)
    compare(ccpu, cgpu, "average_down")
    assert (cgpu.download_state(0) != before).any()
    cgpu.close()


@pytest.mark.parametrize("rank,periodic", [(2, False), (3, False), (2, True), (3, True)])
def test_flux_register_cycle(oracle, fvs, rank, periodic):
    """One coarse step with two fine subcycles, the order of
    SubcycleFineFirstSolver::AdvanceLevel: reset, fine fluxes accumulated with
    FineAdd, coarse fluxes with CrseInit, Reflux, average_down, complete."""
    ccpu, fcpu, cgpu, fgpu, cf = two_levels(oracle, fvs, rank, periodic)
    ccpu.fill_ghost_layer()
    cgpu.fill_ghost_layer()
    cf.fill_ghost_layer_two_levels()
    cgpu.fill_ghost_layer_two_levels(1, 0)
    dt = 0.2 * min(min(ccpu.compute_stable_dt(d), fcpu.compute_stable_dt(d)) for d in range(rank))
    cf.reset()
    cgpu.reset_coarse_fine_fluxes(1, 0)
    for sub in range(2):
        for d in range(rank):
            fcpu.compute_numeric_fluxes(0.5 * dt, d)
            fgpu.compute_numeric_fluxes(0.5 * dt, d)
            fcpu.update_conservatively(0.5 * dt, d)
            fgpu.update_conservatively(0.5 * dt, d)
            fcpu.complete_from_cons()
            fgpu.complete_from_cons()
            cf.accumulate(fcpu, 0.5 * dt, d)
            cgpu.accumulate_coarse_fine_fluxes(1, 0.5 * dt, d)
    for d in range(rank):
        ccpu.compute_numeric_fluxes(dt, d)
        cgpu.compute_numeric_fluxes(dt, d)
        ccpu.update_conservatively(dt, d)
        cgpu.update_conservatively(dt, d)
        ccpu.complete_from_cons()
        cgpu.complete_from_cons()
        cf.accumulate(ccpu, dt, d)
        cgpu.accumulate_coarse_fine_fluxes(0, dt, d)
    before = cgpu.download_state(0).copy()
    cf.apply_flux_correction()
    cgpu.apply_flux_correction(1, 0, dt)
    compare(ccpu, cgpu, "reflux")
    assert (cgpu.download_state(0) != before).any(), "the flux correction changed nothing"
    cf.coarsen_conservatively()
    cgpu.coarsen_conservatively(1, 0)
    ccpu.complete_from_cons()
    cgpu.complete_from_cons()
    compare(ccpu, cgpu, "after average_down + complete")
    compare(fcpu, fgpu, "fine level")
    cgpu.close()


def nested_hierarchy(orc, fvs, nlevels):
    """2-D diagonal Sod problem (examples/AMReX/SodShockTube2d.cpp:57-64) on a
    static, properly nested hierarchy; reflective boundaries, gcw 8/6."""
    from helpers import complete_from_prim
    rank = 2
    n0 = [32, 24]
    level_boxes = [
        [([0, 0, 0], [15, 23, 0]), ([16, 0, 0], [31, 23, 0])],
        [([16, 8, 0], [31, 39, 0]), ([32, 8, 0], [47, 23, 0]), ([32, 24, 0], [47, 39, 0])],
        [([48, 40, 0], [71, 63, 0]), ([72, 32, 0], [87, 55, 0])],
    ][:nlevels]
    gcw = (8, 6)
    cpus, gpus = [], []
    ctx = None
    for l, boxes in enumerate(level_boxes):
        f = 2 ** l
        hi = [n0[0] * f - 1, n0[1] * f - 1, 0]
        dx = [2.0 / (n0[0] * f), 2.0 / (n0[1] * f), 1.0]
        cpu = orc.Level(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
        if l == 0:
            ctx = fvs.Context(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
            gpu = ctx
        else:
            gpu = ctx.define_level(l, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
        for lv in (cpu, gpu):
            for d in range(rank):
                for s in (0, 1):
                    lv.add_bc(d, s, 1)
        for b in range(len(boxes)):
            arr, lo, hi_b = cpu.scratch(b)
            x = -1.0 + (np.arange(lo[0], hi_b[0] + 1) + 0.5) * dx[0]
            y = -1.0 + (np.arange(lo[1], hi_b[1] + 1) + 0.5) * dx[1]
            upper = (x[None, None, :] + y[None, :, None]) > 0.0
            rho = np.where(upper, 0.125, 1.0)
            p = np.where(upper, 10132.5, 101325.0)
            arr[...] = complete_from_prim(rank, rho, [np.zeros_like(rho)] * 2, p)
            gpu.upload_state(b, arr)
        cpus.append(cpu)
        gpus.append(gpu)
    hier = orc.Hierarchy(cpus)
    cpus[0].fill_ghost_layer()
    gpus[0].fill_ghost_layer()
    for l in range(1, nlevels):
        orc.CoarseFine(cpus[l], cpus[l - 1]).fill_ghost_layer_two_levels()
        ctx.fill_ghost_layer_two_levels(l, l - 1)
    return hier, cpus, gpus, ctx


@pytest.mark.parametrize("nlevels,split", [(2, 1), (3, 1), (3, 0)])
def test_subcycled_hierarchy_bitwise(oracle, fvs, nlevels, split):
    """SubcycleFineFirstSolver on a static hierarchy: fine levels first with two
    subcycles each, flux registers, reflux, average_down, two-level ghost fill;
    every level bit-identical, identical per-sweep dts and step dt."""
    hier, cpus, gpus, ctx = nested_hierarchy(oracle, fvs, nlevels)
    for step in range(4):
        dt_c = hier.compute_stable_dt()
        dt_g = ctx.compute_stable_dt_hierarchy()
        assert dt_c == dt_g, (step, dt_c, dt_g)
        ok_c, _, dts_c = hier.advance_hierarchy(0.5 * dt_c, split)
        ok_g, _, dts_g = ctx.advance_hierarchy(0.5 * dt_g, split)
        assert ok_c and ok_g and dts_c == dts_g, (step, dts_c, dts_g)
        assert len(dts_c) == (4 if split == 1 else 2) * (2 ** nlevels - 1)
        for l in range(nlevels):
            compare(cpus[l], gpus[l], f"step {step} level {l}")
    ctx.close()


def test_hierarchy_cfl_violation_rolls_back_all_levels(oracle, fvs):
    hier, cpus, gpus, ctx = nested_hierarchy(oracle, fvs, 2)
    before = [[g.download_state(b) for b in range(len(g.boxes))] for g in gpus]
    dt = 3.0 * ctx.compute_stable_dt_hierarchy()
    ok_c, fail_c, _ = hier.advance_hierarchy(dt, 1)
    ok_g, fail_g, _ = ctx.advance_hierarchy(dt, 1)
    assert not ok_c and not ok_g and fail_c == fail_g
    for l, g in enumerate(gpus):
        for b in range(len(g.boxes)):
            assert_bitwise_equal(g.download_state(b), before[l][b], f"rollback level {l} box {b}")
        compare(cpus[l], g, f"rollback vs oracle level {l}")
    ctx.close()
