"""CPU tests of the oracle's cut-cell and tagging restatement (no GPU).

The reference has no numeric test for these functions ("parity unpinned",
DESIGN.md section 8); the checks here are the algebraic identities the
reference's formulas imply and hand-computed known answers for
FillCutCellData / IsCutCell / TagBuffer."""
import numpy as np
import pytest

from helpers import (complete_from_prim, eb_fab_type, eb_geometry, polygon_phi, sphere_phi)


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
