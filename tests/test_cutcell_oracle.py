"""CPU tests of the oracle's cut-cell and tagging restatement (no GPU).

The reference has no numeric test for these functions ("parity unpinned",
DESIGN.md section 8); the checks here are the algebraic identities the
reference's formulas imply and hand-computed known answers for
FillCutCellData / IsCutCell / TagBuffer."""
import numpy as np
import pytest

from helpers import (NOZZLE_CASE, PLENUM_CASE, complete_from_prim, eb_fab_type, eb_geometry,
                     eb_shock_pair, eb_two_level_wedge, nozzle_phi, plenum_phi, polygon_phi,
                     sphere_phi)


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def wedge_level(orc, n=(40, 28), gcw=(4, 2), method=(0, 1, 2)):
    dx = [0.15 / n[0], 0.10 / n[1]]
    L = orc.Level(2, [0, 0, 0], [n[0] - 1, n[1] - 1, 0], dx, [([0, 0, 0], [n[0] - 1, n[1] - 1, 0])],
                  scratch_gcw=gcw[0], flux_gcw=gcw[1])
    L.set_method(*method)
    L.enable_eb()
    arr, lo, hi = L.scratch(0)
    phi = polygon_phi([(0.02, 0.05), (0.05, 0.0655), (0.05, 0.0345)])
    g = eb_geometry(2, lo, hi, dx + [1.0], [0, 0, 0], phi)
    L.set_eb(0, g, eb_fab_type(g))
    for d, s in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        L.add_bc(d, s, orc.BC_TRANSMISSIVE)
    x = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    rho = np.ones((1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1))
    p = np.where(x < 0.015, 5 * 101325.0, 101325.0)[None, None, :] + np.zeros_like(rho)
    st = complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, p) * (g["alpha"] > 0)
    arr[...] = st
    L.fill_ghost_layer()
    return L, g


def test_fill_cutcell_data_known_answers(oracle):
    L, g = wedge_level(oracle)
    beta = g["beta"][0][0, 0]
    us = L.eb_array(0, oracle.EB_US, 0)[0]
    sl = L.eb_array(0, oracle.EB_SL, 0)[0]
    sr = L.eb_array(0, oracle.EB_SR, 0)[0]
    ds = L.eb_array(0, oracle.EB_DS, 0)[0]
    bL, bM, bR = beta[:, :-2], beta[:, 1:-1], beta[:, 2:]
    dL = np.maximum(0.0, bM - bL)
    dR = np.maximum(0.0, bM - bR)
    want_us = np.maximum(0.0, bM - np.maximum(dL, dR))
    want_ds = np.minimum(dL, dR)
    assert np.array_equal(us[0, 0][:, 1:-1], want_us)
    assert np.array_equal(ds[0, 0][:, 1:-1], want_ds)
    assert np.array_equal(sl[0, 0][:, 1:-1], dL - want_ds)
    assert np.array_equal(sr[0, 0][:, 1:-1], dR - want_ds)
    safe = np.where(bM > 0, bM, 1.0)
    assert np.array_equal(us[1, 0][:, 1:-1], want_us / safe)
    # the four parts add up to the face fraction
    total = us[0, 0] + sl[0, 0] + sr[0, 0] + ds[0, 0]
    assert np.allclose(total[:, 1:-1], bM, atol=1e-15)


def test_is_cut_cell_definition(oracle):
    L, g = wedge_level(oracle)
    a = g["alpha"][0, 0]
    bx, by = g["beta"][0][0, 0], g["beta"][1][0, 0]
    want = (a > 0) & ((a < 1) | (bx[:, :-1] < 1) | (bx[:, 1:] < 1) | (by[:-1, :] < 1) | (by[1:, :] < 1))
    got = L.is_cut_cell(0)[0].astype(bool)
    assert np.array_equal(got, want)
    assert got.sum() > 10


def test_uniform_state_is_preserved_around_the_body(oracle):
    """A gas at rest with uniform pressure is an exact steady state of the
    Kbn scheme: boundary flux = pressure on the wall, stabilised fluxes
    balance, the cut-cell update adds nothing."""
    L, g = wedge_level(oracle)
    arr, lo, hi = L.scratch(0)
    rho = np.ones(arr.shape[1:])
    st = complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, 101325.0 + np.zeros_like(rho))
    arr[...] = st * (g["alpha"][0] > 0)
    before = arr.copy()
    dt = 0.4 * min(L.compute_stable_dt(0), L.compute_stable_dt(1))
    ok, _, _ = L.advance_hierarchy(dt, oracle.SPLIT_STRANG)
    assert ok
    fluid = np.broadcast_to(g["alpha"] > 0, arr.shape)
    assert np.allclose(arr[fluid], before[fluid], rtol=1e-12, atol=1e-9)
    assert np.all(arr[~fluid] == 0.0)


def test_shock_past_wedge_runs_and_conserves_sign(oracle):
    L, g = wedge_level(oracle)
    arr, _, _ = L.scratch(0)
    for _ in range(30):
        dt = 0.4 * min(L.compute_stable_dt(0), L.compute_stable_dt(1))
        ok, _, dts = L.advance_hierarchy(dt, oracle.SPLIT_STRANG)
        assert ok and len(dts) == 4
    assert not np.isnan(arr).any()
    fluid = g["alpha"][0] > 0
    assert arr[0][fluid].min() > 0.5 and arr[0][fluid].max() < 4.0
    assert np.all(arr[:, ~fluid] == 0.0)
    refs = L.eb_array(0, oracle.EB_REFERENCE)[0]
    cut = L.is_cut_cell(0).astype(bool)
    assert np.all(refs[0][cut] > 0) and np.all(refs[0][~cut] == 0)


def test_tag_buffer_2d_known_answer(oracle):
    t = np.zeros((1, 9, 9), dtype=np.int8)
    t[0, 4, 4] = 1
    oracle.tag_buffer(t, 2, 2)
    want = np.zeros((9, 9), dtype=np.int8)
    want[2:6, 2:6] = 3      # [i - w, i + w)
    want[4, 4] = 2
    assert np.array_equal(t[0], want)


def test_tag_buffer_3d_tags_interior(oracle):
    t = np.zeros((7, 7, 7), dtype=np.int8)
    t[3, 3, 3] = 1
    oracle.tag_buffer(t, 3, 2)
    want = np.zeros((7, 7, 7), dtype=np.int8)
    want[:6, :6, :6] = 2
    want[3, 3, 3] = 1
    assert np.array_equal(t, want)


def test_tag_gradient_and_cutcells(oracle):
    L, g = wedge_level(oracle)
    tags = L.tag_gradient(0, [(0, 1e-2), (4, 0.05)])
    arr, lo, hi = L.scratch(0)
    p = arr[4, 0, 4:-4, 4:-4]
    # the pressure jump sits between x-cells 3 and 4 (x = 0.015)
    cols = np.where(tags[0].any(axis=0))[0]
    assert set([3, 4]).issubset(set(cols.tolist()))
    eb_tags = L.tag_gradient(0, [(4, 0.05)], use_alpha=True)
    assert eb_tags.sum() <= tags.sum()
    cc = L.tag_cutcells(0)
    a = g["alpha"][0, 0, 4:-4, 4:-4]
    assert np.array_equal(cc[0].astype(bool), (a > 0) & (a < 1))


def test_cutcell_3d_sphere_runs(oracle):
    n = 16
    dx = [1.0 / n] * 3
    L = oracle.Level(3, [0] * 3, [n - 1] * 3, dx, [([0] * 3, [n - 1] * 3)], scratch_gcw=4, flux_gcw=2)
    L.set_method(2, 1, 1)  # MUSCL<HLL Einfeldt, MinMod> like Divider.cpp
    L.enable_eb()
    arr, lo, hi = L.scratch(0)
    g = eb_geometry(3, lo, hi, dx, [0, 0, 0], sphere_phi((0.5, 0.5, 0.5), 0.22), nsub=4)
    L.set_eb(0, g, eb_fab_type(g))
    for d in range(3):
        for s in (0, 1):
            L.add_bc(d, s, oracle.BC_TRANSMISSIVE)
    x = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    shape = arr.shape[1:]
    rho = np.ones(shape) * 1.22
    p = np.where(x < 0.15, 8 * 101325.0, 101325.0)[None, None, :] + np.zeros(shape)
    arr[...] = complete_from_prim(3, rho, [np.zeros_like(rho)] * 3, p) * (g["alpha"][0] > 0)
    L.fill_ghost_layer()
    for _ in range(4):
        dt = 0.4 * min(L.compute_stable_dt(d) for d in range(3))
        ok, _, _ = L.advance_hierarchy(dt, oracle.SPLIT_GODUNOV)
        assert ok
    assert not np.isnan(arr).any()
    assert L.is_cut_cell(0).sum() > 50


def _eb_composite(C_, F_, geoms, g=8, nc=(48, 32)):
    covered = np.zeros((nc[1], nc[0]), dtype=bool)
    for lo, hi in F_.boxes:
        covered[lo[1] // 2:hi[1] // 2 + 1, lo[0] // 2:hi[0] // 2 + 1] = True
    tot = np.zeros(2)
    for lvl, L, vol in ((0, C_, (0.15 / 48) * (0.10 / 32)), (1, F_, (0.15 / 96) * (0.10 / 64))):
        for b, (lo, hi) in enumerate(L.boxes):
            arr, _, _ = L.scratch(b)
            w = geoms[lvl][b]["alpha"][0, 0, g:-g, g:-g] * vol
            if lvl == 0:
                w = w * ~covered[lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
            tot += np.array([(arr[0, 0, g:-g, g:-g] * w).sum(), (arr[3, 0, g:-g, g:-g] * w).sum()])
    return tot


def _eb_start(orc, **kw):
    cpus, _, _, geoms = eb_two_level_wedge(orc, **kw)
    hier = orc.Hierarchy(cpus)
    cf = orc.CoarseFine(cpus[1], cpus[0])
    cf.coarsen_conservatively()
    cpus[0].complete_from_cons()
    cpus[0].fill_ghost_layer()
    cf.fill_ghost_layer_two_levels()
    return cpus, geoms, hier, cf


def test_eb_two_level_composite_conservation(oracle):
    """Cut-cell levels inside a hierarchy (src/AMReX/cutcell/IntegratorContext.cpp:596-747):
    pc_interp ghost fill, FineAdd / CrseInit / Reflux of the unshielded fluxes and
    EB_average_down through the cut-cell adapters.  With no body in the domain (every fab
    regular, gas at rest at the domain boundary) the composite mass and energy are conserved
    to rounding over subcycled steps.  (With a body they are not, on one level either: the
    sub-sampled stand-in geometry is not closed to rounding.)"""
    orc = oracle
    (C_, F_), geoms, hier, cf = _eb_start(orc, phi=lambda x, y, z: 1.0 + 0 * (x + y + z),
                                          x_shock=0.06)
    before = _eb_composite(C_, F_, geoms)
    for step in range(2):
        dt = 0.4 * hier.compute_stable_dt()
        ok, _, dts = hier.advance_hierarchy(dt, orc.SPLIT_STRANG)
        assert ok and len(dts) == 12
    after = _eb_composite(C_, F_, geoms)
    assert np.max(np.abs(after - before) / np.abs(before)) < 1e-13, (before, after)
    arr, _, _ = F_.scratch(1)
    assert np.ptp(arr[4]) > 1e5  # the shock sits inside the fine level


def test_eb_two_level_wedge_steady_state_and_average_down(oracle):
    """Body inside the fine level.  (1) A gas at rest with uniform pressure stays at rest on
    both levels through subcycled steps -- the plain average_down would mix the zero state of
    covered fine cells into the coarse cut cells, the volume-fraction weighted EB_average_down
    does not.  (2) After a shock run the coarse cells under the fine level hold
    sum(alpha U) / sum(alpha) of their children, the lower-left child where all are covered."""
    orc = oracle
    (C_, F_), geoms, hier, cf = _eb_start(orc, p_ratio=1.0)
    ref = [[L.scratch(b)[0].copy() for b in range(len(L.boxes))] for L in (C_, F_)]
    for step in range(2):
        ok, _, _ = hier.advance_hierarchy(0.4 * hier.compute_stable_dt(), orc.SPLIT_STRANG)
        assert ok
    for lvl, L in enumerate((C_, F_)):
        for b in range(len(L.boxes)):
            arr = L.scratch(b)[0]
            fluid = np.broadcast_to(geoms[lvl][b]["alpha"] > 0, arr.shape)
            assert np.allclose(arr[fluid], ref[lvl][b][fluid], rtol=1e-12, atol=1e-9), (lvl, b)
    (C_, F_), geoms, hier, cf = _eb_start(orc)
    for step in range(3):
        ok, _, _ = hier.advance_hierarchy(0.4 * hier.compute_stable_dt(), orc.SPLIT_STRANG)
        assert ok
    cf.coarsen_conservatively()
    checked = 0
    for fb, (flo, fhi) in enumerate(F_.boxes):
        farr, fl, _ = F_.scratch(fb)
        fa = geoms[1][fb]["alpha"][0, 0]
        for cb, (clo, chi) in enumerate(C_.boxes):
            carr, cl, _ = C_.scratch(cb)
            for j in range(max(flo[1] // 2, clo[1]), min(fhi[1] // 2, chi[1]) + 1):
                for i in range(max(flo[0] // 2, clo[0]), min(fhi[0] // 2, chi[0]) + 1):
                    fj, fi = 2 * j - fl[1], 2 * i - fl[0]
                    al = fa[fj:fj + 2, fi:fi + 2]
                    u = farr[3, 0, fj:fj + 2, fi:fi + 2]
                    got = carr[3, 0, j - cl[1], i - cl[0]]
                    if al.sum() > 1e-30:
                        assert abs(got - (u * al).sum() / al.sum()) <= 4e-16 * abs(got)
                        checked += bool((al < 1).any())
                    else:
                        assert got == u[0, 0]
    assert checked > 10


def test_nozzle_and_plenum_level_sets():
    """The geometries of BASELINE configs 4 and 5 (examples/AMReX/EB/2D/ConvergentNozzle.cpp:164-185,
    examples/AMReX/EB/3D/MultiTube_PerfectGas.cpp:48-83) as level sets: known points."""
    noz = nozzle_phi()
    z = 0.0
    assert noz(0.05, 0.3, z) > 0          # plenum half-space
    assert noz(-0.5, 0.0, z) > 0          # inside the tube
    assert noz(-0.5, 0.0149, z) > 0 and noz(-0.5, 0.0151, z) < 0      # tube wall at r_tube
    assert noz(-1e-4, 0.0074, z) > 0 and noz(-1e-4, 0.0076, z) < 0    # throat at r_tube / 2
    assert noz(-0.03, 0.0112, z) > 0 and noz(-0.03, 0.0113, z) < 0    # half way down the convergent part
    assert noz(-2.5, 0.0, z) < 0          # behind the inlet
    ple = plenum_phi()
    r_c = 0.5 * (0.065 + 0.1925)
    assert ple(0.5, 0.13, 0.0) > 0        # annulus
    assert ple(0.5, 0.05, 0.0) < 0        # hub
    assert ple(0.5, 0.2, 0.0) < 0         # outside the outer wall
    assert ple(0.9, 0.05, 0.0) > 0        # behind the hub (it ends at x = 0.75)
    for i in range(6):                    # the six tube stubs
        a = i * np.pi / 3
        assert ple(-0.5, r_c * np.cos(a), r_c * np.sin(a)) > 0
        assert ple(-0.5, (r_c + 0.016) * np.cos(a), (r_c + 0.016) * np.sin(a)) < 0
    assert ple(-0.5, 0.0, 0.0) < 0        # between the tubes


@pytest.mark.parametrize("case,phi,split", [(NOZZLE_CASE, nozzle_phi(), 1), (PLENUM_CASE, plenum_phi(), 0)])
def test_nozzle_and_plenum_runs(oracle, case, phi, split):
    """Shock running through the inlet into the plenum on the oracle: cut cells exist on both
    boxes' geometries, the run stays finite and physical, covered cells stay zero."""
    cpu, _, geoms = eb_shock_pair(oracle, None, phi=phi, **case)
    rank = case["rank"]
    cut = sum(int(((g["alpha"] > 0) & (g["alpha"] < 1)).sum()) for g in geoms)
    assert cut > (100 if rank == 2 else 1000)
    for _ in range(6):
        dt = 0.4 * min(cpu.compute_stable_dt(d) for d in range(rank))
        ok, _, _ = cpu.advance_hierarchy(dt, split)
        assert ok
    gw = case["gcw"][0]
    inner = (slice(gw, -gw) if rank == 3 else slice(None), slice(gw, -gw), slice(gw, -gw))
    for b, g in enumerate(geoms):
        arr = cpu.scratch(b)[0]
        assert not np.isnan(arr).any()
        fluid = (g["alpha"][0] > 0)[inner]     # valid cells (ghost cells may copy covered ones)
        rho = arr[0][inner]
        assert rho[fluid].min() > 0.3 and rho[fluid].max() < 4.0
        assert np.all(arr[:, g["alpha"][0] == 0] == 0.0)
