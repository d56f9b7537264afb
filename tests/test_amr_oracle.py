"""CPU test of the oracle's coarse-fine restatement (oracle/amr.hpp): on a
periodic domain the composite grid must conserve mass, momentum and energy to
rounding after FineAdd / CrseInit / Reflux / average_down -- which pins the
sign and scaling conventions of the flux register."""
import numpy as np
import pytest

from helpers import smooth_state


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.mark.parametrize("rank,touching", [(2, False), (3, False), (2, True), (3, True)])
def test_composite_conservation(oracle, rank, touching):
    """touching: fine boxes on the periodic domain boundary -- the outside cell of their
    register faces is the coarse cell on the OTHER side of the domain
    (amrex::FluxRegister::Reflux copies with geom.periodicity())."""
    orc = oracle
    if rank == 2:
        nc = [32, 24, 1]
        cboxes = [([0, 0, 0], [15, 23, 0]), ([16, 0, 0], [31, 23, 0])]
        fboxes = [([16, 8, 0], [31, 31, 0]), ([32, 8, 0], [47, 23, 0]), ([32, 24, 0], [47, 39, 0])]
        if touching:  # x low and x high boundary (periodic neighbours of each other), y low
            fboxes = [([0, 8, 0], [15, 31, 0]), ([48, 8, 0], [63, 23, 0]), ([24, 0, 0], [39, 15, 0])]
    else:
        nc = [16, 12, 12]
        cboxes = [([0, 0, 0], [15, 11, 11])]
        fboxes = [([8, 4, 4], [19, 15, 15]), ([20, 4, 4], [27, 15, 19])]
        if touching:
            fboxes = [([0, 4, 4], [11, 15, 15]), ([20, 4, 16], [31, 15, 23])]
    dxc = [1.0 / nc[d] if d < rank else 1.0 for d in range(3)]
    dxf = [0.5 * dxc[d] if d < rank else 1.0 for d in range(3)]
    chi = [nc[d] - 1 if d < rank else 0 for d in range(3)]
    fhi = [2 * nc[d] - 1 if d < rank else 0 for d in range(3)]
    per = [1 if d < rank else 0 for d in range(3)]
    C = orc.Level(rank, [0] * 3, chi, dxc, cboxes, scratch_gcw=4, flux_gcw=2, periodic=per)
    F = orc.Level(rank, [0] * 3, fhi, dxf, fboxes, scratch_gcw=4, flux_gcw=2, periodic=per)
    for b in range(len(cboxes)):
        arr, lo, hi = C.scratch(b)
        arr[...] = smooth_state(rank, lo, hi, dxc, seed=1)
    cf = orc.CoarseFine(F, C)
    # consistent start: fine data = injected smooth field, coarse under fine = its average
    for b in range(len(fboxes)):
        arr, lo, hi = F.scratch(b)
        arr[...] = smooth_state(rank, lo, hi, dxf, seed=1)
    cf.coarsen_conservatively()
    C.complete_from_cons()
    C.fill_ghost_layer()
    cf.fill_ghost_layer_two_levels()

    def composite():
        covered = np.zeros((nc[2], nc[1], nc[0]), dtype=bool)
        for lo, hi in fboxes:
            covered[lo[2] // (2 if rank > 2 else 1):hi[2] // (2 if rank > 2 else 1) + 1,
                    lo[1] // 2:hi[1] // 2 + 1, lo[0] // 2:hi[0] // 2 + 1] = True
        tot = np.zeros(rank + 2)
        volc = np.prod(dxc[:rank])
        volf = np.prod(dxf[:rank])
        for b, (lo, hi) in enumerate(cboxes):
            arr, slo, _ = C.scratch(b)
            g = 4
            sl = (slice(None), slice(g if rank > 2 else 0, -g if rank > 2 else None),
                  slice(g, -g), slice(g, -g))
            v = arr[sl][:rank + 2]
            m = ~covered[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
            tot += (v * m).sum(axis=(1, 2, 3)) * volc
        for b in range(len(fboxes)):
            arr, _, _ = F.scratch(b)
            g = 4
            sl = (slice(None), slice(g if rank > 2 else 0, -g if rank > 2 else None),
                  slice(g, -g), slice(g, -g))
            tot += arr[sl][:rank + 2].sum(axis=(1, 2, 3)) * volf
        return tot

    before = composite()
    dt = 0.3 * min(min(C.compute_stable_dt(d), F.compute_stable_dt(d)) for d in range(rank))
    cf.reset()
    for sub in range(2):
        for d in range(rank):
            F.compute_numeric_fluxes(0.5 * dt, d)
            F.update_conservatively(0.5 * dt, d)
            F.complete_from_cons()
            cf.accumulate(F, 0.5 * dt, d)
    for d in range(rank):
        C.compute_numeric_fluxes(dt, d)
        C.update_conservatively(dt, d)
        C.complete_from_cons()
        cf.accumulate(C, dt, d)
    no_reflux = composite()
    cf.apply_flux_correction()
    cf.coarsen_conservatively()
    C.complete_from_cons()
    after = composite()
    scale = np.abs(before) + 1e-3
    # without the flux correction the composite grid is NOT conservative ...
    assert np.max(np.abs(no_reflux - before) / scale) > 1e-9
    # ... with it, it is, to rounding
    assert np.max(np.abs(after - before) / scale) < 1e-13, (before, after)
