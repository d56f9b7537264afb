"""Parity at BASELINE.json's full size (configs[1]: 3-D spherical Riemann
problem on 512^3, MUSCL-Hancock<HLLEM-L, VanLeer>, Strang splitting, gcw 8/6).

The oracle cannot run 512^3 in seconds, so the check uses two size-independent
properties of the scheme:
 * locality: with the reference's wide-halo cadence a box only needs its ghost
   cells at step entry, so a 40^3 sub-box of the big domain, run on the oracle
   with ghost cells taken from the same initial data, must reproduce the GPU's
   512^3 result in its valid cells BIT FOR BIT after one coarse step;
 * conservation: with transmissive boundaries and a wave that has not reached
   them, total mass and energy are conserved to rounding."""
import numpy as np
import pytest

from helpers import assert_bitwise_equal, riemann_state

pytestmark = pytest.mark.gpu

N = 512
GCW = (8, 6)


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


def test_512_cubed_step_subbox_bitwise_and_conservation(oracle, fvs):
    dx = [2.0 / N] * 3
    origin = [-1.0] * 3
    ctx = fvs.Context(3, [0] * 3, [N - 1] * 3, dx, [([0] * 3, [N - 1] * 3)], scratch_gcw=GCW[0],
                      flux_gcw=GCW[1])
    for d in range(3):
        for s in (0, 1):
            ctx.add_bc(d, s, fvs.BC_TRANSMISSIVE)
    lo, hi, _ = ctx.state_box(0)
    state = np.empty(ctx.state_shape(0))
    for k in range(lo[2], hi[2] + 1):  # plane by plane: bounded host memory
        state[:, k - lo[2]] = riemann_state(3, [lo[0], lo[1], k], [hi[0], hi[1], k], dx, origin)[:, 0]
    ctx.upload_state(0, state)
    ctx.fill_ghost_layer()
    g = GCW[0]
    valid = (slice(None), slice(g, -g), slice(g, -g), slice(g, -g))
    before = state[valid][[0, 4]].sum(axis=(1, 2, 3))
    dt = 0.5 * ctx.compute_stable_dt_all()
    ok, _, dts = ctx.advance_hierarchy(dt, fvs.SPLIT_STRANG)
    assert ok and len(dts) == 6
    got = ctx.download_state(0)
    ctx.close()
    assert np.isfinite(got).all()
    after = got[valid][[0, 4]].sum(axis=(1, 2, 3))
    assert np.max(np.abs(after - before) / before) < 1e-12

    # sub-boxes: across the discontinuity (r = 0.4 -> cell 256 +- 102), and at rest
    for sub_lo in ([338, 250, 250], [300, 310, 230], [100, 100, 100]):
        sub_hi = [v + 39 for v in sub_lo]
        lvl = oracle.Level(3, [0] * 3, [N - 1] * 3, dx, [(sub_lo, sub_hi)], scratch_gcw=GCW[0],
                           flux_gcw=GCW[1])
        lvl.set_method(oracle.BASE_HLLEM_LARROUTUROU, oracle.RECON_MUSCL_HANCOCK,
                       oracle.LIMITER_VANLEER)
        arr, slo, shi = lvl.scratch(0)
        arr[...] = riemann_state(3, slo, shi, dx, origin)
        ok_c, _, _ = lvl.advance_hierarchy(dt, oracle.SPLIT_STRANG)
        assert ok_c
        sl = (slice(None),) + tuple(slice(sub_lo[d] - lo[d], sub_hi[d] - lo[d] + 1)
                                    for d in (2, 1, 0))
        ref = arr[(slice(None),) + (slice(g, -g),) * 3]
        assert_bitwise_equal(got[sl], ref, f"512^3 step, sub-box {sub_lo}")
        if sub_lo[0] != 100:
            assert (ref[1:4] != 0).any(), "the sub-box should see the wave"
