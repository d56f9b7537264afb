"""Pins the CPU oracle against the reference's own unit tests.

Each test names the reference test it restates (paths relative to
/root/reference).  These run without a GPU.
"""
import math

import numpy as np
import pytest

GAMMA = 1.4


def complete(orc, rank, rho, mom, energy):
    return orc.complete_from_cons(rank, [rho, *mom, energy])


# --- tests/test.HyperbolicSplitPatchIntegrator.cpp:37-102 --------------------
def test_update_1_cell(oracle):
    out = oracle.update_conservatively([1.0], [0.3, 0.4], 1.0, 1.0)
    assert out[0] == pytest.approx(0.9)
    # exactly what prev + dt/dx (fL - fR) gives in IEEE arithmetic
    assert out[0] == 1.0 + 1.0 * (0.3 - 0.4)


def test_update_8_cells(oracle):
    prev = np.ones(8)
    flux = np.array([0.3, 0.4, -0.2, 0.0, 0.1, 0.2, 0.0, 0.0, 0.0])
    out = oracle.update_conservatively(prev, flux, 1.0, 1.0)
    for i in range(8):
        assert out[i] == pytest.approx(prev[i] + flux[i] - flux[i + 1])
        assert out[i] == prev[i] + 1.0 * (flux[i] - flux[i + 1])


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 15, 16, 17, 64])
def test_update_widths(oracle, n):
    # tests/benchmark/HyperbolicSplitPatchIntegrator.cpp:14-76 widths, tails
    rng = np.random.default_rng(1234)
    prev = rng.random(n)
    flux = rng.random(n + 1)
    out = oracle.update_conservatively(prev, flux, 0.37, 0.0125)
    lam = 0.37 / 0.0125
    ref = prev + lam * (flux[:-1] - flux[1:])
    assert np.array_equal(out, ref)


# --- tests/test.simd_FluxMethod.cpp:54-105 -----------------------------------
def test_simd_single_state_muscl_hll(oracle):
    orc = oracle
    s0 = complete(orc, 2, 1.0, [0.0, 0.0], 2.0)
    s2 = complete(orc, 2, 1.0, [0.0, 0.0], 1.0)
    states = np.concatenate([s0, s0, s2, s2])
    # MusclHancockMethod(equation, hll): default limiter MinMod
    scalar = orc.numeric_flux_scalar_hll(2, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_MINMOD,
                                         states, 1.0, 1.0, 0)
    array = orc.numeric_flux(2, orc.BASE_HLL_EINFELDT, orc.RECON_MUSCL_HANCOCK,
                             orc.LIMITER_MINMOD, states, 1.0, 1.0, 0)
    assert np.array_equal(scalar, array)
    assert np.all(np.isfinite(array))
    # physical sanity: equal densities at rest -> no mass flux; the pressure
    # jump gives positive x-momentum and energy flux
    assert array[0] == 0.0 and array[1] > 0.0 and array[3] > 0.0


# --- tests/test.simd_FluxMethod.cpp:40-52,107-186 ------------------------------
def _init_states(orc, nx, ny, gx, gy, dx):
    # InitializeStates: rho=1, m=0, E = 1 + 0.5 sin(x^2 + y^2); x = i dx
    st = {}
    for j in range(-gy, ny + gy):
        for i in range(-gx, nx + gx):
            x, y = i * dx, j * dx
            st[(i, j)] = complete(orc, 2, 1.0, [0.0, 0.0], 1.0 + 0.5 * math.sin(x * x + y * y))
    return st


@pytest.mark.parametrize("direction", [0, 1])
def test_hll_patch_scalar_equals_simd(oracle, direction):
    orc = oracle
    dx = 0.0125
    gx, gy = (1, 0) if direction == 0 else (0, 1)
    st = _init_states(orc, 16, 16, gx, gy, dx)
    # array path through the level machinery (16x16 box, 1 ghost in `direction`)
    lvl = orc.Level(2, [0, 0], [15, 15], [dx, dx], [([0, 0], [15, 15])],
                    scratch_gcw=1, flux_gcw=0)
    lvl.set_method(orc.BASE_HLL_EINFELDT, orc.RECON_FIRST_ORDER, orc.LIMITER_NONE)
    arr, lo, hi = lvl.scratch(0)
    for j in range(lo[1], hi[1] + 1):
        for i in range(lo[0], hi[0] + 1):
            x, y = i * dx, j * dx
            arr[:, 0, j - lo[1], i - lo[0]] = complete(
                orc, 2, 1.0, [0.0, 0.0], 1.0 + 0.5 * math.sin(x * x + y * y))
    dt = lvl.compute_stable_dt(direction)
    assert 0.0 < dt < 1.0
    lvl.compute_numeric_fluxes(dt, direction)
    flux, flo, fhi = lvl.fluxes(0, direction)
    # scalar path face by face
    for j in range(0, 16 + (direction == 1)):
        for i in range(0, 16 + (direction == 0)):
            left = st[(i - 1, j)] if direction == 0 else st[(i, j - 1)]
            right = st[(i, j)]
            scalar = orc.numeric_flux_scalar_hll(2, orc.RECON_FIRST_ORDER, orc.LIMITER_NONE,
                                                 np.concatenate([left, right]), dt, dx,
                                                 direction)
            got = flux[:, 0, j - flo[1], i - flo[0]]
            assert np.array_equal(scalar, got), (i, j)


# --- tests/test.PerfectGas.ExactRiemannProblem.cpp:55-135 + Toro Table 4.2/4.3 -
def _prim_state(rho, u, p):
    # the test's from_prim (energy = p/(gamma-1); momentum given directly)
    return [rho, rho * u, p * (1.0 / (GAMMA - 1.0)), p, math.sqrt(GAMMA * p / rho)]


def _toro_pstar(rhoL, uL, pL, rhoR, uR, pR):
    """Independent textbook solver (Toro ch. 4): bisection on f(p) = 0."""
    g = GAMMA
    aL, aR = math.sqrt(g * pL / rhoL), math.sqrt(g * pR / rhoR)

    def fk(p, rho, pk, ak):
        if p > pk:
            A, B = 2.0 / ((g + 1) * rho), (g - 1) / (g + 1) * pk
            return (p - pk) * math.sqrt(A / (p + B))
        return 2 * ak / (g - 1) * ((p / pk) ** ((g - 1) / (2 * g)) - 1.0)

    def f(p):
        return fk(p, rhoL, pL, aL) + fk(p, rhoR, pR, aR) + (uR - uL)

    lo, hi = 1e-12, 10 * max(pL, pR) + 10
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0:
            hi = mid
        else:
            lo = mid
    p = 0.5 * (lo + hi)
    u = 0.5 * (uL + uR) + 0.5 * (fk(p, rhoR, pR, aR) - fk(p, rhoL, pL, aL))
    return p, u


TORO = [
    # (rhoL, uL, pL, rhoR, uR, pR, p*, u*) -- Toro Table 4.3 where available
    (1.0, 0.0, 1.0, 0.125, 0.0, 0.1, 0.30313, 0.92745),
    (1.0, -2.0, 0.4, 1.0, 2.0, 0.4, 0.00189, 0.0),
    (1.0, 0.0, 1000.0, 1.0, 0.0, 0.01, 460.894, 19.5975),
    (1.0, 0.0, 0.1, 1.0, 0.0, 100.0, None, None),  # reference uses pL=0.1 (:124)
    (1.0, 0.0, 0.01, 1.0, 0.0, 100.0, 46.0950, -6.19633),  # Toro test 4 proper
]


@pytest.mark.parametrize("case", TORO)
def test_exact_riemann_toro(oracle, case):
    rhoL, uL, pL, rhoR, uR, pR, pstar, ustar = case
    pu, sol = oracle.exact_riemann(1, _prim_state(rhoL, uL, pL), _prim_state(rhoR, uR, pR))
    p_ref, u_ref = _toro_pstar(rhoL, uL, pL, rhoR, uR, pR)
    # Newton tolerance in the reference is |f| < 1e-7
    assert pu[0] == pytest.approx(p_ref, rel=1e-5, abs=1e-6)
    assert pu[1] == pytest.approx(u_ref, rel=1e-5, abs=1e-6)
    if pstar is not None:
        assert pu[0] == pytest.approx(pstar, rel=2e-4, abs=2e-5)
        assert pu[1] == pytest.approx(ustar, rel=2e-4, abs=2e-5)
    assert np.all(np.isfinite(sol))
    assert sol[0] > 0 and sol[3] > 0


def test_exact_riemann_sod_sample(oracle):
    # Sod: x/t = 0 lies inside the left rarefaction fan (sonic point not reached),
    # textbook star-left state: rho*L = 0.42632, u* = 0.92745, p* = 0.30313
    pu, sol = oracle.exact_riemann(1, _prim_state(1.0, 0.0, 1.0), _prim_state(0.125, 0.0, 0.1))
    assert sol[0] == pytest.approx(0.42632, rel=2e-4)
    assert sol[1] / sol[0] == pytest.approx(0.92745, rel=2e-4)
    assert sol[3] == pytest.approx(0.30313, rel=2e-4)


# --- tests/test.State.cpp:40-66: SoA layout rho, m[3], E, p, a ----------------
def test_state_layout(oracle):
    lvl = oracle.Level(3, [0, 0, 0], [7, 7, 7], [1, 1, 1], [([0, 0, 0], [7, 7, 7])],
                       scratch_gcw=0, flux_gcw=0)
    arr, lo, hi = lvl.scratch(0)
    assert arr.shape == (7, 8, 8, 8)
    assert arr.strides == (8 * 8 * 8 * 8, 8 * 8 * 8, 8 * 8, 8)  # x fastest, comp slowest


# --- equation constants (PerfectGas.hpp:216-219) -----------------------------
def test_gamma_minus_1_inv_is_not_2_5(oracle):
    gm1inv = 1.0 / (1.4 - 1.0)
    assert gm1inv != 2.5
    q = oracle.complete_from_cons(1, [1.0, 0.0, 2.5])
    assert q[3] == 2.5 / gm1inv
    assert q[4] == math.sqrt(1.4 * q[3] / 1.0)


def test_kinetic_energy_reduction_order(oracle):
    # Eigen unrolled redux: m0^2 + (m1^2 + m2^2)
    m = [0.1, 0.7, 1.3]
    q = oracle.complete_from_cons(3, [1.7, *m, 9.0])
    ke = 0.5 * (m[0] * m[0] + (m[1] * m[1] + m[2] * m[2])) / 1.7
    assert q[5] == (9.0 - ke) / (1.0 / (1.4 - 1.0))


import pytest as _pytest


@_pytest.mark.parametrize("recon", [2, 3])
def test_muscl_hancock_prim_and_char_sod_star_state(oracle, recon):
    """perfect_gas::MusclHancockPrim (2) and MusclHancockCharacteristic (3): the
    reference has no test for them; the oracle's restatement must resolve the
    Sod problem -- star density 0.42632, post-contact density 0.26557 (Toro,
    section 4.3.3) at the right positions."""
    import numpy as np
    from helpers import complete_from_prim
    n = 400
    dx = [2.0 / n]
    lvl = oracle.Level(1, [0], [n - 1], dx, [([0], [n - 1])], scratch_gcw=2, flux_gcw=0)
    lvl.set_method(oracle.BASE_HLLEM_LARROUTUROU, recon, 2)
    lvl.add_bc(0, 0, oracle.BC_TRANSMISSIVE)
    lvl.add_bc(0, 1, oracle.BC_TRANSMISSIVE)
    arr, lo, hi = lvl.scratch(0)
    x = -1.0 + (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    rho = np.where(x < 0.0, 1.0, 0.125)[None, None, :]
    p = np.where(x < 0.0, 1.0, 0.1)[None, None, :]
    arr[...] = complete_from_prim(1, rho, [np.zeros_like(rho)], p)
    lvl.fill_ghost_layer()
    lvl.run(0.5, 0.25, -1, oracle.SPLIT_GODUNOV)
    r = arr[0, 0, 0]
    assert not np.isnan(arr).any()
    star = r[(x > 0.08) & (x < 0.20)]           # between contact (0.232) and rarefaction tail
    assert abs(star.mean() - 0.42632) < 3e-3
    post_shock = r[(x > 0.27) & (x < 0.41)]     # between contact and shock (0.438)
    assert abs(post_shock.mean() - 0.26557) < 3e-3
    assert abs(r[x > 0.5].mean() - 0.125) < 1e-12


@_pytest.mark.parametrize("recon", [4, 5, 6])
@_pytest.mark.parametrize("limiter", [2, 3])
@_pytest.mark.parametrize("base", [0, 1, 2])
def test_muscl_hancock2_sod_star_state(oracle, recon, limiter, base):
    """MusclHancock2<Gradient, Reconstruction, Base> (MusclHancockMethod2.hpp; upstream only
    constructs it, tests/test.NewStates.cpp): conservative / primitive / characteristic
    variables with the MinMod and VanLeer limiters of Gradient.hpp must resolve the Sod
    problem (Toro, section 4.3.3: star density 0.42632, post-contact density 0.26557)."""
    import numpy as np
    from helpers import complete_from_prim
    n = 400
    dx = [2.0 / n]
    lvl = oracle.Level(1, [0], [n - 1], dx, [([0], [n - 1])], scratch_gcw=2, flux_gcw=0)
    lvl.set_method(base, recon, limiter)
    lvl.add_bc(0, 0, oracle.BC_TRANSMISSIVE)
    lvl.add_bc(0, 1, oracle.BC_TRANSMISSIVE)
    arr, lo, hi = lvl.scratch(0)
    x = -1.0 + (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    rho = np.where(x < 0.0, 1.0, 0.125)[None, None, :]
    p = np.where(x < 0.0, 1.0, 0.1)[None, None, :]
    arr[...] = complete_from_prim(1, rho, [np.zeros_like(rho)], p)
    lvl.fill_ghost_layer()
    lvl.run(0.5, 0.25, -1, oracle.SPLIT_GODUNOV)
    r = arr[0, 0, 0]
    assert not np.isnan(arr).any()
    star = r[(x > 0.08) & (x < 0.20)]
    assert abs(star.mean() - 0.42632) < 3e-3
    post_shock = r[(x > 0.27) & (x < 0.41)]
    assert abs(post_shock.mean() - 0.26557) < 3e-3
    assert abs(r[x > 0.5].mean() - 0.125) < 1e-12


@_pytest.mark.parametrize("recon", [4, 5, 6])
def test_muscl_hancock2_limiters_on_known_stencils(oracle, recon):
    """Gradient.hpp:31-59, :211-217 through the flux of a four-cell stencil: (a) a uniform
    stencil gives the physical flux for every limiter; (b) the upwind limiter (zero gradient)
    gives exactly the flux of the base method on the two inner cells; (c) across an extremum
    (sL sR <= 0 in every variable) every limiter agrees with upwind."""
    import numpy as np
    rank = 1
    q = lambda rho, u, p: complete(oracle, rank, rho, [rho * u], p / (GAMMA - 1.0) + 0.5 * rho * u * u)
    dt, dx = 0.01, 0.1
    uniform = np.concatenate([q(1.3, 0.4, 0.9)] * 4)
    f_phys = np.array([1.3 * 0.4, 1.3 * 0.4 * 0.4 + 0.9,
                       0.4 * (0.9 / (GAMMA - 1.0) + 0.5 * 1.3 * 0.16 + 0.9)])
    for lim in range(4):
        f = oracle.numeric_flux(rank, 0, recon, lim, uniform, dt, dx, 0)
        assert np.allclose(f, f_phys, rtol=1e-14, atol=0)
    st = np.concatenate([q(1.0, 0.1, 1.0), q(1.2, 0.3, 1.4), q(0.9, 0.2, 0.8), q(1.1, 0.5, 1.0)])
    first_order = oracle.numeric_flux(rank, 0, 0, 0, st[5:15], dt, dx, 0)
    upwind = oracle.numeric_flux(rank, 0, recon, 1, st, dt, dx, 0)
    assert np.allclose(upwind, first_order, rtol=1e-14, atol=0)
    if recon == 4:
        # zig-zag in rho, rho u and E alike: every central gradient is masked out
        zig = np.concatenate([q(1.0, 0.0, 1.0), q(2.0, 0.0, 2.0), q(1.0, 0.0, 1.0), q(2.0, 0.0, 2.0)])
        ref = oracle.numeric_flux(rank, 0, recon, 1, zig, dt, dx, 0)
        for lim in (0, 2, 3):
            assert np.array_equal(oracle.numeric_flux(rank, 0, recon, lim, zig, dt, dx, 0), ref)


# --- HLLEM: the reference's two code paths against each other ---------------------
# The reference cross-checks its scalar and array paths for HLL only
# (tests/test.simd_FluxMethod.cpp:54-105).  oracle/scalar_hllem.hpp restates the
# scalar HLLEM path (src/equations/perfect_gas/HllemMethod.hpp:371-640)
# independently of the array restatement the product is compared with
# (oracle/numerics.hpp <- :711-1072); the two differ upstream in the order of the
# Roe averages (equal up to rounding) and in the Larrouturou upwinding of a flux
# towards the left (quirk Q6, scalar_hllem.hpp header).
def _random_states(orc, rank, rng, supersonic=False):
    out = []
    for _ in range(2):
        rho = 0.2 + 2.0 * rng.random()
        p = 0.1 + 3.0 * rng.random()
        a = math.sqrt(GAMMA * p / rho)
        u = (rng.random(rank) - 0.5) * (6.0 if supersonic else 1.0) * a
        e = p / (GAMMA - 1.0) + 0.5 * rho * float(np.dot(u, u))
        out.append(complete(orc, rank, rho, list(rho * u), e))
    return np.concatenate(out)


@pytest.mark.parametrize("rank", [1, 2])
@pytest.mark.parametrize("supersonic", [False, True])
def test_hllem_scalar_path_agrees_with_array_path(oracle, rank, supersonic):
    orc = oracle
    rng = np.random.default_rng(17 + rank)
    worst = 0.0
    for _ in range(400):
        states = _random_states(orc, rank, rng, supersonic)
        for direction in range(rank):
            scalar = orc.hllem_scalar(rank, False, states, direction)
            array = orc.numeric_flux(rank, orc.BASE_HLLEM, orc.RECON_FIRST_ORDER,
                                     orc.LIMITER_NONE, states, 1.0, 1.0, direction)
            scale = np.max(np.abs(array)) + 1e-300
            worst = max(worst, float(np.max(np.abs(scalar - array)) / scale))
    # only the order of the Roe averages differs: a few ulp, amplified by the
    # cancellation in roeH - |roeU|^2 / 2
    assert worst < 2e-14, worst


def test_hllem_scalar_equals_array_bitwise_for_equal_densities(oracle):
    """With rhoL == rhoR both Roe weights are exactly 1/2 and every product with them is
    exact, so the two operation orders round identically: bit-for-bit equality of two
    independent restatements of the same formulas."""
    orc = oracle
    rng = np.random.default_rng(5)
    for rank in (1, 2):
        for _ in range(300):
            rho = 4.0 ** rng.integers(-2, 3)   # sqrt(rho) exact
            cells = []
            for _ in range(2):
                p = 0.1 + 3.0 * rng.random()
                u = (rng.random(rank) - 0.5)
                e = p / (GAMMA - 1.0) + 0.5 * rho * float(np.dot(u, u))
                cells.append(complete(orc, rank, rho, list(rho * u), e))
            states = np.concatenate(cells)
            for direction in range(rank):
                scalar = orc.hllem_scalar(rank, False, states, direction)
                array = orc.numeric_flux(rank, orc.BASE_HLLEM, orc.RECON_FIRST_ORDER,
                                         orc.LIMITER_NONE, states, 1.0, 1.0, direction)
                assert np.array_equal(scalar, array), (rank, direction, scalar, array)


def test_hllem_larrouturou_scalar_vs_array_upwinding(oracle):
    """Q6: for a mass flux towards the right the scalar and the array path agree (to the
    rounding of the Roe averages); towards the left the scalar path carries MINUS the
    upwinded tangential momentum flux (HllemMethod.hpp:543-546 vs :907-910)."""
    orc = oracle
    rng = np.random.default_rng(23)
    seen = {True: 0, False: 0}
    for _ in range(600):
        states = _random_states(orc, 2, rng)
        for direction in (0, 1):
            iy = 1 - direction
            scalar = orc.hllem_scalar(2, True, states, direction)
            array = orc.numeric_flux(2, orc.BASE_HLLEM_LARROUTUROU, orc.RECON_FIRST_ORDER,
                                     orc.LIMITER_NONE, states, 1.0, 1.0, direction)
            scale = np.max(np.abs(array)) + 1e-300
            assert abs(scalar[0] - array[0]) / scale < 2e-14
            assert abs(scalar[1 + direction] - array[1 + direction]) / scale < 2e-14
            towards_right = array[0] > 0.0
            seen[towards_right] += 1
            if towards_right:
                assert np.max(np.abs(scalar - array)) / scale < 2e-14
            else:
                assert abs(scalar[1 + iy] + array[1 + iy]) / scale < 2e-14
    assert seen[True] > 100 and seen[False] > 100


# --- known answers of the HLLEM flux and of the limiter ---------------------------
@pytest.mark.parametrize("base", [0, 1])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_hllem_keeps_a_stationary_contact_exactly(oracle, rank, base):
    """The point of the HLLEM anti-diffusion (Einfeldt 1991): across a contact at rest
    (jump in density, equal pressure, zero velocity) the numeric flux is the physical
    one, (0, p e_d, 0) -- no mass or energy flux, where plain HLLE smears the contact.
    Powers of four keep sqrt(rho) and the Roe weights exact, so 'exactly' means bits."""
    orc = oracle
    p = 1.0
    left = complete(orc, rank, 4.0, [0.0] * rank, p / (GAMMA - 1.0))
    right = complete(orc, rank, 1.0, [0.0] * rank, p / (GAMMA - 1.0))
    for direction in range(rank):
        f = orc.numeric_flux(rank, base, orc.RECON_FIRST_ORDER, orc.LIMITER_NONE,
                             np.concatenate([left, right]), 1.0, 1.0, direction)
        hll = orc.numeric_flux(rank, orc.BASE_HLL_EINFELDT, orc.RECON_FIRST_ORDER,
                               orc.LIMITER_NONE, np.concatenate([left, right]), 1.0, 1.0, direction)
        assert abs(f[0]) < 1e-15 and abs(f[rank + 1]) < 1e-15, f
        assert abs(f[1 + direction] - p) < 1e-15
        assert hll[0] > 0.1, "plain HLLE diffuses the contact (mass flux towards the light gas)"


def test_q1_3d_normal_momentum_antidiffusion_vanishes(oracle):
    """Quirk Q1 (HllemMethod.hpp:991-992): the array 3-D path subtracts
    b_delta_alpha_2 * roeU.row(iy) from the NORMAL momentum flux.  With the Larrouturou
    modification roeU[iy] is zero, so in 3-D the normal momentum flux of HLLEM-L equals
    the HLLE value, while the same states in 2-D (which uses row(ix)) get the correction."""
    orc = oracle
    rng = np.random.default_rng(3)
    differs_2d = 0
    for _ in range(50):
        rho = [0.5 + rng.random(), 0.5 + rng.random()]
        p = [0.5 + rng.random(), 0.5 + rng.random()]
        u = [0.3 * (rng.random() - 0.5), 0.3 * (rng.random() - 0.5)]
        def cells(rank):
            out = []
            for k in range(2):
                vel = [u[k]] + [0.0] * (rank - 1)
                e = p[k] / (GAMMA - 1.0) + 0.5 * rho[k] * u[k] * u[k]
                out.append(complete(orc, rank, rho[k], [rho[k] * v for v in vel], e))
            return np.concatenate(out)
        f3 = orc.numeric_flux(3, orc.BASE_HLLEM_LARROUTUROU, orc.RECON_FIRST_ORDER,
                              orc.LIMITER_NONE, cells(3), 1.0, 1.0, 0)
        # the HLLE value of the normal momentum flux with HLLEM's signal speeds (:752-811)
        a = [math.sqrt(GAMMA * p[k] / rho[k]) for k in range(2)]
        en = [p[k] / (GAMMA - 1.0) + 0.5 * rho[k] * u[k] * u[k] for k in range(2)]
        h = [(en[k] + p[k]) / rho[k] for k in range(2)]
        sq = [math.sqrt(r) for r in rho]
        w = [sq[k] / (sq[0] + sq[1]) for k in range(2)]
        roe_u = w[0] * u[0] + w[1] * u[1]
        roe_a = math.sqrt((GAMMA - 1.0) * (w[0] * h[0] + w[1] * h[1] - 0.5 * roe_u * roe_u))
        b_l = min(min(u[0] - max(a), roe_u - roe_a), 0.0)
        b_r = max(max(roe_u + roe_a, u[1] + max(a)), 0.0)
        m = [rho[k] * u[k] for k in range(2)]
        fm = [u[k] * m[k] + p[k] for k in range(2)]
        h3 = [None, (b_r * fm[0] - b_l * fm[1] + b_l * b_r * (m[1] - m[0])) / (b_r - b_l)]
        f2 = orc.numeric_flux(2, orc.BASE_HLLEM_LARROUTUROU, orc.RECON_FIRST_ORDER,
                              orc.LIMITER_NONE, cells(2), 1.0, 1.0, 0)
        f1 = orc.numeric_flux(1, orc.BASE_HLLEM_LARROUTUROU, orc.RECON_FIRST_ORDER,
                              orc.LIMITER_NONE, cells(1), 1.0, 1.0, 0)
        assert abs(f3[1] - h3[1]) <= 1e-14 * abs(h3[1])
        # mass and energy flux are the same in every dimension; the normal momentum
        # flux of 1-D and 2-D carries the anti-diffusion term that 3-D loses
        assert abs(f3[0] - f2[0]) <= 1e-15 and abs(f2[0] - f1[0]) <= 1e-15
        assert abs(f2[1] - f1[1]) <= 1e-15
        differs_2d += abs(f2[1] - f3[1]) > 1e-6
    assert differs_2d > 40


def test_vanleer_and_minmod_known_values(oracle):
    """MusclHancockMethod.hpp:92-187: VanLeer(sL, sR) = 2 sL sR / (sL + sR) for sL sR > 0,
    else 0; MinMod picks the smaller slope.  Read off the reconstructed density: with a
    uniform velocity-free state only rho varies, q0 - q_L(edge) = slope / 2 ... probed through
    the first-order-in-time limit dt -> 0 of MUSCL-Hancock over HLL: the flux through the
    face between cells 1 and 2 is then the HLL flux of (q1 + s1/2, q2 - s2/2)."""
    orc = oracle
    p = 1.0

    def cell(rho):
        return complete(orc, 1, rho, [0.0], p / (GAMMA - 1.0))

    def face_flux(limiter, rhos, dt=0.0):
        states = np.concatenate([cell(r) for r in rhos])
        return orc.numeric_flux(1, orc.BASE_HLL_EINFELDT, orc.RECON_MUSCL_HANCOCK, limiter,
                                states, dt, 1.0, 0)

    def hll(rho_l, rho_r):
        return orc.numeric_flux(1, orc.BASE_HLL_EINFELDT, orc.RECON_FIRST_ORDER, orc.LIMITER_NONE,
                                np.concatenate([cell(rho_l), cell(rho_r)]), 0.0, 1.0, 0)

    # densities 1, 2, 5, 7: slopes of cell 1 are (1, 3), of cell 2 (3, 2)
    rhos = [1.0, 2.0, 5.0, 7.0]
    # VanLeer(1, 3) = 1.5, VanLeer(3, 2) = 2.4 -> edge states 2.75 | 3.8
    assert np.array_equal(face_flux(orc.LIMITER_VANLEER, rhos), hll(2.0 + 0.75, 5.0 - 1.2))
    # MinMod(1, 3) = 1, MinMod(3, 2) = 2 -> 2.5 | 4
    assert np.array_equal(face_flux(orc.LIMITER_MINMOD, rhos), hll(2.5, 4.0))
    # NoLimiter: central slopes (1 + 3) / 2 = 2, (3 + 2) / 2 = 2.5 -> 3 | 3.75
    assert np.array_equal(face_flux(orc.LIMITER_NONE, rhos), hll(3.0, 3.75))
    # an extremum: slopes of opposite sign -> VanLeer and MinMod give 0 (first order)
    rhos = [1.0, 3.0, 2.0, 4.0]
    assert np.array_equal(face_flux(orc.LIMITER_VANLEER, rhos), hll(3.0, 2.0))
    assert np.array_equal(face_flux(orc.LIMITER_MINMOD, rhos), hll(3.0, 2.0))
