"""The opt-in tolerance arithmetic of the fused sweep (fvs_set_arithmetic,
FVS_ARITH_TOLERANCE) against the CPU oracle.

The default arithmetic is bit-identical to the oracle (test_gpu_parity.py).  This
mode trades the last ulp for throughput (reciprocal-multiply quotients, FMA
contraction) and must stay inside the tolerance BASELINE.json's north_star states:
conserved variables within 1e-12 relative L-inf per step, 1e-10 after 100 steps;
the dt sequence agrees to the same relative accuracy (it cannot be bit-identical
once the states differ in the last place).
"""
import numpy as np
import pytest

from helpers import Pair, complete_from_prim, rel_linf, riemann_state, smooth_state
from test_gpu_parity import SIZES, geometry, multi_boxes

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12   # north_star: conserved variables per step
TOL_100 = 1e-10    # north_star: after 100 steps

METHODS = [
    # (base, recon, limiter)
    (0, 1, 2),  # MUSCL-Hancock<HLLEM Larrouturou, VanLeer>  (headline)
    (0, 1, 1), (0, 1, 0), (0, 0, 0),
    (1, 1, 2), (1, 0, 0),
    (2, 1, 1), (2, 1, 2), (2, 0, 0),
    (0, 2, 0),  # perfect_gas::MusclHancockPrim
    (0, 3, 2),  # perfect_gas::MusclHancockCharacteristic
]


@pytest.fixture(scope="module")
def fvs():
    from finitevolumesolver_b200 import capi
    assert capi.device_count() > 0, "no CUDA device"
    return capi


def conserved_rel_linf(pair, rank):
    """max over boxes and conserved components of |got - ref|_inf / |ref|_inf"""
    worst = 0.0
    for b in range(pair.nboxes):
        ref, _, _ = pair.cpu.scratch(b)
        got = pair.gpu.download_state(b)
        assert np.isfinite(got).all()
        for c in range(rank + 2):
            worst = max(worst, rel_linf(got[c], ref[c]))
    return worst


@pytest.mark.parametrize("rank", [1, 2, 3])
@pytest.mark.parametrize("method", METHODS)
def test_fused_sweep_within_step_tolerance(oracle, fvs, rank, method):
    """Every sweep starts from the oracle's state (re-uploaded), so the figure
    is the per-sweep error, which bounds the per-step one."""
    base, recon, limiter = method
    n = SIZES[rank]
    lo, hi, dx = geometry(rank, n, 4)
    pair = Pair(oracle, fvs, rank, lo[:rank], hi[:rank], dx[:rank], [(lo[:rank], hi[:rank])],
                4, 2, base=base, recon=recon, limiter=limiter)
    pair.gpu.set_arithmetic(fvs.ARITH_TOLERANCE)
    pair.fill(lambda l, h: smooth_state(rank, l, h, dx, seed=2 + rank))
    differs = False
    for sweep in range(2 * rank):
        d = sweep % rank
        dt_ref = pair.cpu.compute_stable_dt(d)
        dt = 0.45 * dt_ref
        pair.cpu.sweep(dt, d)
        rc, level_dt = pair.gpu.sweep(dt, d)
        assert rc == fvs.OK
        # the stable dt of the (identical) input state
        assert abs(level_dt - dt_ref) <= TOL_STEP * dt_ref, (d, level_dt, dt_ref)
        worst = conserved_rel_linf(pair, rank)
        assert worst <= TOL_STEP, f"sweep {sweep} dir {d}: {worst:.3e}"
        ref, _, _ = pair.cpu.scratch(0)
        got = pair.gpu.download_state(0)
        differs = differs or not np.array_equal(got, ref)
        # p and a are recomputed from the conserved variables
        assert rel_linf(got[rank + 2], ref[rank + 2]) <= 10 * TOL_STEP
        assert rel_linf(got[rank + 3], ref[rank + 3]) <= 10 * TOL_STEP
        pair.gpu.upload_state(0, ref)
    assert differs, "tolerance arithmetic gave bit-identical results: is the build wired?"
    pair.close()


def test_tolerance_arithmetic_is_opt_in_and_reversible(oracle, fvs):
    rank, n = 2, (24, 16)
    lo, hi, dx = geometry(rank, n, 4)
    pair = Pair(oracle, fvs, rank, lo[:2], hi[:2], dx[:2], [(lo[:2], hi[:2])], 4, 2)
    pair.fill(lambda l, h: smooth_state(rank, l, h, dx))
    pair.gpu.set_arithmetic(fvs.ARITH_TOLERANCE)
    pair.gpu.set_arithmetic(fvs.ARITH_IEEE)
    dt = 0.4 * pair.cpu.compute_stable_dt(0)
    pair.cpu.sweep(dt, 0)
    pair.gpu.sweep(dt, 0)
    pair.compare_states("IEEE arithmetic after switching back")  # bitwise
    with pytest.raises(RuntimeError):
        pair.gpu.set_arithmetic(7)
    pair.gpu.set_method(3, 1, 1)  # exact Riemann solver: no tolerance build
    with pytest.raises(RuntimeError):
        pair.gpu.set_arithmetic(fvs.ARITH_TOLERANCE)
    pair.close()


def test_degenerate_states_fall_back_to_ieee(oracle, fvs):
    """A zero state and a denormal density leave the reciprocal's domain; the
    stage is then recomputed with the plain operators, exactly as the default
    arithmetic does.  Compared against the default arithmetic on the device:
    same status, unaffected cells within tolerance."""
    rank, n = 1, (64,)
    lo, hi, dx = geometry(rank, n, 2)
    out = []
    for mode in (fvs.ARITH_IEEE, fvs.ARITH_TOLERANCE):
        pair = Pair(oracle, fvs, rank, lo[:1], hi[:1], dx[:1], [(lo[:1], hi[:1])], 2, 0)
        pair.gpu.set_arithmetic(mode)

        def fill(l, h):
            s = smooth_state(rank, l, h, dx, seed=1)
            s[:, :, :, 20] = 0.0          # a zero state: divisions by rho = 0
            s[0, :, :, 40] = 1e-310       # denormal density
            return s

        pair.fill(fill)
        rc, level_dt = pair.gpu.sweep(1e-9, 0)
        out.append((rc, level_dt, pair.gpu.download_state(0)))
        pair.close()
    (rc_i, dt_i, ref), (rc_t, dt_t, got) = out
    assert rc_t == rc_i
    assert np.isnan(dt_t) == np.isnan(dt_i)
    if np.isfinite(dt_i):
        assert abs(dt_t - dt_i) <= 1e-9 * abs(dt_i)
    # cells whose stencil saw a degenerate state hold inf / NaN / huge values
    # (which FMA contraction may round differently); everything else agrees
    far = np.ones(ref.shape[-1], dtype=bool)
    for c in (20, 40):
        far[max(c - 4, 0):c + 5] = False
    assert np.isfinite(got[..., far]).all() and np.isfinite(ref[..., far]).all()
    assert rel_linf(got[..., far], ref[..., far]) <= TOL_STEP


def test_sod_1d_run_within_tolerance(oracle, fvs):
    """BASELINE config 1 (examples/AMReX/PerfectGas1d.cpp) to t = 0.25."""
    n = 200
    dx = [2.0 / n]
    bcs = [(0, 0, 1), (0, 1, 1)]
    pair = Pair(oracle, fvs, 1, [0], [n - 1], dx, [([0], [n - 1])], 2, 0, bcs=bcs)
    pair.gpu.set_arithmetic(fvs.ARITH_TOLERANCE)

    def sod(l, h):
        x = -1.0 + (np.arange(l[0], h[0] + 1) + 0.5) * dx[0]
        rho = np.where(x < 0.0, 1.0, 0.125)[None, None, :]
        p = np.where(x < 0.0, 1.0, 0.1)[None, None, :]
        return complete_from_prim(1, rho, [np.zeros_like(rho)], p)

    pair.fill(sod)
    pair.cpu.fill_ghost_layer()
    pair.gpu.fill_ghost_layer()
    dts_c, retries_c = pair.cpu.run(0.5, 0.25, -1, 0)
    dts_g, retries_g = pair.gpu.run(0.5, 0.25, -1, 0)
    assert retries_c == retries_g
    assert len(dts_c) == len(dts_g) and len(dts_c) > 50
    assert np.max(np.abs(np.array(dts_g) - np.array(dts_c)) / np.array(dts_c)) <= TOL_100
    assert conserved_rel_linf(pair, 1) <= TOL_100
    pair.close()


def test_spherical_riemann_3d_100_steps_within_tolerance(oracle, fvs):
    """BASELINE config 2 at oracle-sized resolution: 32^3, Strang, gcw 8/6,
    transmissive, CFL 0.5, 100 steps: 1e-12 after the first step, 1e-10 after
    100, dt sequence to 1e-12 relative."""
    rank, n = 3, (32, 32, 32)
    lo, hi, _ = geometry(rank, n, 8)
    dx = [2.0 / 32] * 3
    bcs = [(d, s, 0) for d in range(3) for s in (0, 1)]
    pair = Pair(oracle, fvs, rank, lo, hi, dx, multi_boxes(3, n, (2, 2, 2)), 8, 6, bcs=bcs)
    pair.gpu.set_arithmetic(fvs.ARITH_TOLERANCE)
    pair.fill(lambda l, h: riemann_state(rank, l, h, dx, [-1.0] * 3))
    pair.cpu.fill_ghost_layer()
    pair.gpu.fill_ghost_layer()
    dts_c, _ = pair.cpu.run(0.5, 1e9, 1, 1)
    dts_g, _ = pair.gpu.run(0.5, 1e9, 1, 1)
    assert abs(dts_g[0] - dts_c[0]) <= TOL_STEP * dts_c[0]  # from identical states
    first = conserved_rel_linf(pair, rank)
    assert first <= TOL_STEP, f"after one Strang step: {first:.3e}"
    more_c, _ = pair.cpu.run(0.5, 1e9, 100, 1)  # max_cycles counts from the start
    more_g, _ = pair.gpu.run(0.5, 1e9, 100, 1)
    assert len(more_c) == len(more_g) == 99
    rel_dt = np.max(np.abs(np.array(more_g) - np.array(more_c)) / np.array(more_c))
    assert rel_dt <= TOL_STEP, f"dt sequence: {rel_dt:.3e}"
    worst = conserved_rel_linf(pair, rank)
    assert worst <= TOL_100, f"after 100 steps: {worst:.3e}"
    print(f"tolerance arithmetic: 1 step {first:.2e}, 100 steps {worst:.2e}, dt {rel_dt:.2e}")
    pair.close()
