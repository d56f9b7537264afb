"""States that leave the domain of the branch-free divisions and square roots
(csrc/fastmath.cuh): the fused sweeps accumulate the validity of their fast
paths (the y/z sweep over a whole line segment, which it then marches again with
the checked stages; the x sweep per stage) and must still return the reference's
values bit for bit -- whatever they are -- in every cell, and the same dt.

Covered: momenta that are tiny but not zero (below 2^-940: the remainder of the
quotient is inexact), denormal momenta with a zero high word, a gas at rest
(+-0 numerators stay on the fast path), near-vacuum and very dense cells, huge
and tiny pressures, a cold cell (internal energy ~ 0).  NaN-producing cells are
left out: NaN payloads differ between the x86 oracle and the device.
"""
import numpy as np
import pytest

from helpers import Pair, bits, complete_from_prim, smooth_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fvs():
    from finitevolumesolver_b200 import capi
    assert capi.device_count() > 0, "no CUDA device"
    return capi


def make_pair(oracle, fvs, rank, n, recon=1, limiter=2, base=0):
    lo = [0] * rank
    hi = [n[d] - 1 for d in range(rank)]
    dx = [1.0 / n[d] if d < rank else 1.0 for d in range(3)]
    bcs = [(d, s, 0) for d in range(rank) for s in (0, 1)]
    pair = Pair(oracle, fvs, rank, lo, hi, dx[:rank], [(lo, hi)], 4, 2, base=base, recon=recon,
                limiter=limiter, bcs=bcs)
    return pair, dx


KINDS = ["tiny_momentum", "denormal_momentum", "at_rest", "thin_cell", "dense_cell",
         "hot_cell", "cold_cell", "tiny_everything", "mixed"]


def extreme_state(rank, dx, kind, seed):
    def fn(l, h):
        rng = np.random.default_rng(seed)
        s = smooth_state(rank, l, h, dx, seed=5)
        rho = s[0].copy()
        vel = [s[1 + d] / s[0] for d in range(rank)]
        p = s[rank + 2].copy()
        shape = rho.shape

        def some_cells(k):
            return tuple(rng.integers(0, m, size=k) for m in shape)

        if kind == "at_rest":
            for d in range(rank):
                vel[d][...] = 0.0
            vel[0][some_cells(6)] = -0.0
        if kind == "tiny_everything":
            rho = rho * 1e-120
            p = p * 1e-120
        if kind in ("thin_cell", "mixed"):
            rho[some_cells(3)] = 1e-30
        if kind in ("dense_cell", "mixed"):
            rho[some_cells(3)] = 1e25
        if kind in ("hot_cell", "mixed"):
            p[some_cells(3)] = 1e40
        if kind == "cold_cell":
            p[some_cells(3)] = 1e-200
        out = complete_from_prim(rank, rho, vel, p)
        if kind in ("tiny_momentum", "mixed"):
            idx = some_cells(8)
            for d in range(rank):
                out[1 + d][idx] = rng.choice([1e-300, -3e-290, 7e-285, 0.0], size=8)
        if kind == "denormal_momentum":
            idx = some_cells(8)
            for d in range(rank):
                out[1 + d][idx] = rng.choice([4.9e-324, -2e-320, 1e-310, 0.0], size=8)
        return out
    return fn


def compare(pair, what):
    """Bit for bit; a NaN equals a NaN (the payloads of x86 and the device differ)."""
    ref, _, _ = pair.cpu.scratch(0)
    got = pair.gpu.download_state(0)
    neq = (bits(got) != bits(ref)) & ~(np.isnan(got) & np.isnan(ref))
    assert not neq.any(), (what, int(neq.sum()), np.argwhere(neq)[0], got[neq][0], ref[neq][0])


@pytest.mark.parametrize("rank,n", [(2, (40, 36)), (3, (36, 20, 24))])
@pytest.mark.parametrize("recon", [1, 2])  # MUSCL-Hancock (shared-memory pipeline), MusclHancockPrim
@pytest.mark.parametrize("kind", KINDS)
def test_fused_sweeps_bitwise_on_states_outside_the_fast_path_domain(oracle, fvs, rank, n, recon,
                                                                     kind):
    pair, dx = make_pair(oracle, fvs, rank, n, recon=recon)
    pair.gpu.set_dt_hints(0)
    pair.fill(extreme_state(rank, dx, kind, seed=17))
    has_nan = False
    with np.errstate(all="ignore"):
        for d in list(range(rank)) + list(reversed(range(rank))):
            dt_ref = pair.cpu.compute_stable_dt(d)
            # a real step where the field stays smooth, a minute one around the violent cells
            # (their neighbours would otherwise go negative and NaN payloads differ)
            gentle = kind in ("tiny_momentum", "denormal_momentum", "at_rest", "tiny_everything")
            step = (0.3 if gentle else 1e-9) * dt_ref
            pair.cpu.sweep(step, d)
            rc, level_dt = pair.gpu.sweep(step, d)
            assert rc == fvs.OK, (kind, d, rc)
            # with NaN cells in the field the maximum over the signal speeds depends on which
            # operand a comparison drops (implementation defined upstream, test_gpu_lazy_dt.py)
            if not has_nan:
                assert level_dt == dt_ref, (kind, d, level_dt, dt_ref)
            has_nan = bool(np.isnan(pair.cpu.scratch(0)[0]).any())
            compare(pair, f"{kind}, direction {d}")
    pair.close()


@pytest.mark.parametrize("rank,n", [(2, (900, 6)), (2, (6, 300)), (3, (40, 12, 20))])
@pytest.mark.parametrize("kind", ["tiny_everything", "tiny_momentum"])
@pytest.mark.parametrize("cap", [None, 5])
def test_repair_list_overflow(oracle, fvs, rank, n, kind, cap, monkeypatch):
    """Many marked steps per line, and (cap = 5, FVS_REPAIR_CAP) more marked steps than the
    repair list of the sweep holds: the repair pass must then recompute every cell of the
    sweep."""
    if cap is not None:
        monkeypatch.setenv("FVS_REPAIR_CAP", str(cap))
    pair, dx = make_pair(oracle, fvs, rank, n)
    pair.gpu.set_dt_hints(0)

    def fill(l, h):
        out = extreme_state(rank, dx, "tiny_everything", seed=3)(l, h)
        if kind == "tiny_momentum":
            out = extreme_state(rank, dx, "at_rest", seed=3)(l, h)
            out[1] = np.where((np.indices(out[1].shape).sum(axis=0) % 2) == 0, 3e-300, -2e-295)
        return out

    pair.fill(fill)
    with np.errstate(all="ignore"):
        for d in range(rank):
            dt_ref = pair.cpu.compute_stable_dt(d)
            pair.cpu.sweep(0.3 * dt_ref, d)
            rc, level_dt = pair.gpu.sweep(0.3 * dt_ref, d)
            assert rc == fvs.OK and level_dt == dt_ref, (kind, d, rc, level_dt, dt_ref)
            compare(pair, f"{kind}, direction {d}")
    pair.close()
