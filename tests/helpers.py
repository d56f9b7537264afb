"""Shared set-up for the parity tests: builds the same level on the CPU oracle
and on the CUDA path and fills both with identical data."""
import numpy as np

GAMMA = 1.4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bitwise_equal(got, ref, what=""):
    """Bit-for-bit equality (signed zeros included)."""
    got = np.ascontiguousarray(got)
    ref = np.ascontiguousarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    neq = bits(got) != bits(ref)
    if neq.any():
        idx = np.argwhere(neq)[0]
        t = tuple(idx)
        rel = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300))
        raise AssertionError(
            f"{what}: {int(neq.sum())} of {neq.size} values differ, first at {t}: "
            f"got {got[t]!r} ref {ref[t]!r}; max rel diff {rel:.3e}")


def rel_linf(got, ref):
    scale = np.max(np.abs(ref))
    return float(np.max(np.abs(got - ref)) / max(scale, 1e-300))


def complete_from_prim(rank, rho, vel, p):
    """[rank+4, ...] Complete state arrays from primitive fields (same fp64
    operation order as the reference's CompleteFromPrim, PerfectGas.cpp:146-158)."""
    mom = [rho * v for v in vel]
    if rank == 1:
        sq = mom[0] * mom[0]
    elif rank == 2:
        sq = mom[0] * mom[0] + mom[1] * mom[1]
    else:
        sq = mom[0] * mom[0] + (mom[1] * mom[1] + mom[2] * mom[2])
    e_kin = 0.5 * sq / rho
    gm1_inv = 1.0 / (GAMMA - 1.0)
    E = e_kin + p * gm1_inv
    a = np.sqrt(GAMMA * p / rho)
    return np.stack([rho, *mom, E, p, a])


def cell_centers(lo, hi, dx, origin):
    """z, y, x coordinate arrays broadcastable to [z, y, x]."""
    xs = [origin[d] + (np.arange(lo[d], hi[d] + 1) + 0.5) * dx[d] for d in range(3)]
    return (xs[2][:, None, None], xs[1][None, :, None], xs[0][None, None, :])


def smooth_state(rank, lo, hi, dx, seed=0):
    """A smooth, strictly subsonic-ish field with all velocity components
    non-zero; deterministic in the GLOBAL cell index (so neighbouring boxes
    agree in their overlap)."""
    z, y, x = cell_centers(lo, hi, dx, (0.0, 0.0, 0.0))
    shape = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    ph = 0.37 * seed
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * x + ph) * np.cos(2 * np.pi * y) * np.cos(3 * z) + np.zeros(shape)
    u = 0.3 * np.sin(2 * np.pi * (x + y) + ph) + 0.1 * z + np.zeros(shape)
    v = -0.2 * np.cos(2 * np.pi * (x - z)) + 0.15 * y + np.zeros(shape)
    w = 0.25 * np.sin(4 * np.pi * y + ph) * np.cos(2 * np.pi * x) + np.zeros(shape)
    p = 1.0 + 0.3 * np.cos(2 * np.pi * (x + 0.5 * y) + ph) * np.sin(2 * z + 1.0) + np.zeros(shape)
    return complete_from_prim(rank, rho, [u, v, w][:rank], p)


def riemann_state(rank, lo, hi, dx, origin, radius=0.4, inner=(1.0, 1.0), outer=(0.125, 0.1)):
    """Spherical (circular / 1-D) Riemann problem at rest: SURVEY 8d W1/W2."""
    z, y, x = cell_centers(lo, hi, dx, origin)
    shape = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    r2 = x * x + np.zeros(shape)
    if rank >= 2:
        r2 = r2 + y * y
    if rank >= 3:
        r2 = r2 + z * z
    inside = r2 < radius * radius
    rho = np.where(inside, inner[0], outer[0])
    p = np.where(inside, inner[1], outer[1])
    zero = np.zeros(shape)
    return complete_from_prim(rank, rho, [zero] * rank, p)


def eb_geometry(rank, lo, hi, dx, origin, phi, nsub=8):
    """Embedded-boundary arrays for the fab box [lo, hi] from a level-set
    function phi(x, y, z) (> 0 in the fluid), by deterministic sub-sampling:
    alpha = fluid fraction of nsub^rank samples per cell, beta[d] the same on
    the faces, boundary centroid = closest interface point to the cell centre
    (units of dx, clipped to [-1/2, 1/2]), bnorm = normal pointing into the
    body (AMReX convention; the reference negates it).  This stands in for
    AMReX EB2 (absent here): oracle and GPU get the SAME arrays."""
    off = (np.arange(nsub) + 0.5) / nsub
    n = [hi[d] - lo[d] + 1 for d in range(3)]

    def coords(d, nodal, sub):
        base = origin[d] + (np.arange(lo[d], hi[d] + 1 + (1 if nodal else 0))) * dx[d]
        if d >= rank:
            return base[:, None] * 0.0 + 0.0 * off[None, :1]
        if nodal:
            return base[:, None] + 0.0 * off[None, :1]
        return base[:, None] + (off[None, :] if sub else np.array([[0.5]])) * dx[d]

    def fraction(nodal_dir):
        cs = [coords(d, d == nodal_dir, True) for d in range(3)]  # [n_d, s_d]
        X = cs[0][None, None, :, None, None, :]
        Y = cs[1][None, :, None, None, :, None]
        Z = cs[2][:, None, None, :, None, None]
        val = phi(X, Y, Z) > 0.0
        val = np.broadcast_to(val, (cs[2].shape[0], cs[1].shape[0], cs[0].shape[0],
                                    cs[2].shape[1], cs[1].shape[1], cs[0].shape[1]))
        return val.reshape(val.shape[:3] + (-1,)).mean(axis=3)[None]

    alpha = fraction(-1)
    beta = [fraction(d) for d in range(rank)]
    # a face is open only if both of its cells hold fluid (EB2 consistency)
    for d in range(rank):
        ax = 3 - d  # axis of direction d in [1, z, y, x]
        covered = (alpha == 0.0)
        sl_lo = [slice(None)] * 4
        sl_hi = [slice(None)] * 4
        sl_lo[ax] = slice(1, None)   # faces 1..n  have a left cell  0..n-1
        sl_hi[ax] = slice(0, -1)     # faces 0..n-1 have a right cell 0..n-1
        beta[d][tuple(sl_lo)] = np.where(covered, 0.0, beta[d][tuple(sl_lo)])
        beta[d][tuple(sl_hi)] = np.where(covered, 0.0, beta[d][tuple(sl_hi)])
    # a cell with a partly closed face is a cut cell (EB2 consistency: the body that
    # covers part of a face reaches into both of its cells); the volume sampling can miss
    # a thin sliver, so such a "full" cell gets the smallest cut the sampling resolves
    partial = np.zeros(alpha.shape, dtype=bool)
    for d in range(rank):
        ax = 3 - d
        lo_f = [slice(None)] * 4
        hi_f = [slice(None)] * 4
        lo_f[ax] = slice(0, -1)
        hi_f[ax] = slice(1, None)
        partial |= (beta[d][tuple(lo_f)] < 1.0) | (beta[d][tuple(hi_f)] < 1.0)
    alpha = np.where(partial & (alpha == 1.0), 1.0 - 0.5 / nsub ** rank, alpha)
    cc = [coords(d, False, False)[:, 0] for d in range(3)]
    X, Y, Z = cc[0][None, None, :], cc[1][None, :, None], cc[2][:, None, None]
    p0 = phi(X, Y, Z) + np.zeros((n[2], n[1], n[0]))
    h = 1e-4 * dx[0]
    grads = []
    for d in range(rank):
        e = [0.0, 0.0, 0.0]
        e[d] = h
        grads.append((phi(X + e[0], Y + e[1], Z + e[2]) - phi(X - e[0], Y - e[1], Z - e[2])) / (2 * h)
                     + np.zeros_like(p0))
    norm = np.sqrt(sum(g * g for g in grads))
    norm = np.where(norm > 0, norm, 1.0)
    cut = (alpha[0] > 0) & (alpha[0] < 1)
    bnorm = np.stack([np.where(cut, -g / norm, 0.0) for g in grads])
    bcent = np.stack([np.where(cut, np.clip(-p0 * (grads[d] / norm) / norm / dx[d], -0.5, 0.5), 0.0)
                      for d in range(rank)])
    return {"alpha": alpha, "beta": beta, "bcent": bcent, "bnorm": bnorm}


def eb_fab_type(geom):
    a = geom["alpha"]
    if np.all(a == 0.0):
        return -1
    if np.all(a == 1.0) and all(np.all(b == 1.0) for b in geom["beta"]):
        return 0
    return 1


class Pair:
    """The same single-level problem on the oracle and on the GPU."""

    def __init__(self, orc, fvs, rank, domain_lo, domain_hi, dx, boxes, scratch_gcw, flux_gcw,
                 periodic=(0, 0, 0), base=0, recon=1, limiter=2, bcs=(), gamma=GAMMA):
        self.rank = rank
        self.nboxes = len(boxes)
        self.cpu = orc.Level(rank, domain_lo, domain_hi, dx, boxes, scratch_gcw=scratch_gcw,
                             flux_gcw=flux_gcw, periodic=periodic, gamma=gamma)
        self.cpu.set_method(base, recon, limiter)
        self.gpu = fvs.Context(rank, domain_lo, domain_hi, dx, boxes, scratch_gcw=scratch_gcw,
                               flux_gcw=flux_gcw, periodic=periodic, gamma=gamma, base=base,
                               recon=recon, limiter=limiter)
        for (d, s, k) in bcs:
            self.cpu.add_bc(d, s, k)
            self.gpu.add_bc(d, s, k)

    def fill(self, fn):
        """fn(lo, hi) -> [ncomp, z, y, x] state for the fab box [lo, hi]."""
        for b in range(self.nboxes):
            arr, lo, hi = self.cpu.scratch(b)
            data = fn(lo, hi)
            arr[...] = data
            self.gpu.upload_state(b, data)

    def compare_states(self, what="state", tol=None):
        worst = 0.0
        for b in range(self.nboxes):
            ref, _, _ = self.cpu.scratch(b)
            got = self.gpu.download_state(b)
            if tol is None:
                assert_bitwise_equal(got, ref, f"{what} box {b}")
            else:
                worst = max(worst, rel_linf(got, ref))
                assert worst <= tol, f"{what} box {b}: rel Linf {worst:.3e} > {tol:.1e}"
        return worst

    def close(self):
        self.gpu.close()


def polygon_phi(points):
    """Level set (> 0 outside) of a convex polygon: max over its edges of the
    signed distance to the edge line."""
    pts = [tuple(map(float, p)) for p in points]
    cx = sum(p[0] for p in pts) / len(pts)
    cy = sum(p[1] for p in pts) / len(pts)
    edges = []
    for i in range(len(pts)):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % len(pts)]
        nx, ny = (y1 - y0), -(x1 - x0)
        ln = float(np.hypot(nx, ny))
        nx, ny = nx / ln, ny / ln
        if (cx - x0) * nx + (cy - y0) * ny > 0:  # make the normal point outwards
            nx, ny = -nx, -ny
        edges.append((x0, y0, nx, ny))

    def phi(x, y, z):
        v = None
        for (x0, y0, nx, ny) in edges:
            t = (x - x0) * nx + (y - y0) * ny
            v = t if v is None else np.maximum(v, t)
        return v + 0 * z
    return phi


def sphere_phi(center, radius, inside_fluid=False):
    def phi(x, y, z):
        r = np.sqrt((x - center[0]) ** 2 + (y - center[1]) ** 2 + (z - center[2]) ** 2)
        return (radius - r) if inside_fluid else (r - radius)
    return phi


WEDGE = [(0.02, 0.05), (0.05, 0.0655), (0.05, 0.0345)]


def eb_two_level_wedge(orc, fvs=None, n0=(48, 32), gcw=(8, 6), method=(0, 1, 2), at_rest=True,
                       p_ratio=5.0, phi=None, x_shock=0.015):
    """The Schardin wedge (examples/AMReX/EB/2D/Schardin.cpp:87-130) on a static
    two-level cut-cell hierarchy: the fine level (ratio 2) covers the whole body
    plus a margin of regular cells, so every coarse-fine face is a regular face.
    Same geometry arrays (per level, from the same level set) on the oracle and,
    if `fvs` is given, on the device.  Returns (cpu levels, gpu levels, ctx,
    geometries[level][box])."""
    rank = 2
    extent = (0.15, 0.10)
    phi = phi or polygon_phi(WEDGE)
    level_boxes = [
        [([0, 0, 0], [n0[0] // 2 - 1, n0[1] - 1, 0]), ([n0[0] // 2, 0, 0], [n0[0] - 1, n0[1] - 1, 0])],
        # body: x in [0.02, 0.05], y in [0.0345, 0.0655] -> coarse cells [6..16] x [11..20] at 48 x 32
        [([8, 16, 0], [23, 31, 0]), ([24, 16, 0], [39, 31, 0]), ([8, 32, 0], [39, 47, 0])],
    ]
    cpus, gpus, geoms = [], [], []
    ctx = None
    for l, boxes in enumerate(level_boxes):
        f = 2 ** l
        hi = [n0[0] * f - 1, n0[1] * f - 1, 0]
        dx = [extent[0] / (n0[0] * f), extent[1] / (n0[1] * f), 1.0]
        cpu = orc.Level(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
        cpu.set_method(*method)
        cpu.enable_eb()
        gpu = None
        if fvs is not None:
            if l == 0:
                ctx = fvs.Context(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1],
                                  base=method[0], recon=method[1], limiter=method[2])
                gpu = ctx
            else:
                gpu = ctx.define_level(l, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
            gpu.enable_eb()
        for lv in (cpu, gpu):
            if lv is None:
                continue
            for d in range(rank):
                for s in (0, 1):
                    lv.add_bc(d, s, 0)
        lg = []
        for b in range(len(boxes)):
            arr, lo, hi_b = cpu.scratch(b)
            g = eb_geometry(rank, lo, hi_b, dx, [0.0] * 3, phi)
            t = eb_fab_type(g)
            cpu.set_eb(b, g, t)
            if gpu is not None:
                gpu.set_eb(b, g, t)
            x = (np.arange(lo[0], hi_b[0] + 1) + 0.5) * dx[0]
            y = (np.arange(lo[1], hi_b[1] + 1) + 0.5) * dx[1]
            shape = arr.shape[1:]
            rho = 1.0 + np.zeros(shape)
            if not at_rest:
                rho = rho + 0.05 * np.sin(40 * x)[None, None, :] * np.cos(30 * y)[None, :, None]
            p = np.where(x < x_shock, p_ratio * 101325.0, 101325.0)[None, None, :] + np.zeros(shape)
            vel = [np.zeros(shape), np.zeros(shape)]
            if not at_rest:
                vel = [5.0 * np.cos(25 * y)[None, :, None] + np.zeros(shape),
                       -3.0 * np.sin(35 * x)[None, None, :] + np.zeros(shape)]
            st = complete_from_prim(rank, rho, vel, p) * (g["alpha"][0] > 0)
            arr[...] = st
            if gpu is not None:
                gpu.upload_state(b, st)
            lg.append(g)
        cpus.append(cpu)
        gpus.append(gpu)
        geoms.append(lg)
    return cpus, gpus, ctx, geoms


def nozzle_phi(r_tube=0.015, height=2.0, center=(0.0, 0.0)):
    """Level set (> 0 in the fluid) of examples/AMReX/EB/2D/ConvergentNozzle.cpp:164-185:
    the plenum half-space x > center_x united with the inlet tube, a polygon of half-height
    r_tube that converges to r_tube / 2 over the last 4 r_tube before the plenum.  (AMReX
    states it as the intersection of the two BODIES, PlaneIF(..., false) and the inverted
    polygon.)"""
    cx, cy = center
    xlo, xhi = cx - height, cx
    r, r2 = r_tube, 0.5 * r_tube
    xdiv = xhi - 4.0 * r
    outside = polygon_phi([(xlo, cy + r), (xdiv, cy + r), (xhi, cy + r2), (xhi, cy - r2),
                           (xdiv, cy - r), (xlo, cy - r)])

    def phi(x, y, z):
        return np.maximum(x - cx + 0 * y, -outside(x, y, z)) + 0 * z
    return phi


def plenum_phi(r_tube=0.015, r_inner=0.065, r_outer=0.1925, n_tubes=6):
    """Level set (> 0 in the fluid) of the annular plenum with its tube stubs,
    examples/AMReX/EB/3D/MultiTube_PerfectGas.cpp:48-83 (.py:25-30): the outer cylinder
    (axis x, 0 <= x <= 1) united with n_tubes tubes of radius r_tube on the circle of radius
    (r_inner + r_outer) / 2 (-2.05 <= x <= 0.05), minus the inner cylinder (-0.25 <= x <= 0.75)."""
    r_c = 0.5 * r_inner + 0.5 * r_outer
    alpha = 2.0 * np.pi / n_tubes

    def cyl(radius, length, cx, cy, cz):
        def f(x, y, z):
            rho = np.sqrt((y - cy) ** 2 + (z - cz) ** 2)
            return np.minimum(radius - rho, 0.5 * length - np.abs(x - cx))
        return f
    outer = cyl(r_outer, 1.0, 0.5, 0.0, 0.0)
    inner = cyl(r_inner, 1.0, 0.25, 0.0, 0.0)
    tubes = [cyl(r_tube, 2.1, -1.0, r_c * np.cos(i * alpha), r_c * np.sin(i * alpha))
             for i in range(n_tubes)]

    def phi(x, y, z):
        fluid = outer(x, y, z)
        for t in tubes:
            fluid = np.maximum(fluid, t(x, y, z))
        return np.minimum(fluid, -inner(x, y, z))
    return phi


def eb_shock_pair(orc, fvs, rank, n, origin, extent, phi, method, gcw, boxes=None, shock_x=0.0,
                  p_ratio=5.0, rho0=1.22, nsub=4, riemann=0):
    """A single cut-cell level with the geometry `phi` on the oracle and (if fvs is given) on the
    device: gas at rest, pressure p_ratio * 101325 left of shock_x.  Returns (cpu, gpu, geoms)."""
    dx = [extent[d] / n[d] for d in range(rank)] + [1.0] * (3 - rank)
    hi = [n[d] - 1 if d < rank else 0 for d in range(3)]
    boxes = boxes or [([0] * 3, hi)]
    cpu = orc.Level(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1])
    cpu.set_method(*method)
    cpu.enable_eb(riemann)
    gpu = None
    if fvs is not None:
        gpu = fvs.Context(rank, [0] * 3, hi, dx, boxes, scratch_gcw=gcw[0], flux_gcw=gcw[1],
                          base=method[0], recon=method[1], limiter=method[2])
        gpu.enable_eb(riemann)
    for lv in (cpu, gpu):
        if lv is not None:
            for d in range(rank):
                for s in (0, 1):
                    lv.add_bc(d, s, 0)
    geoms = []
    for b in range(len(boxes)):
        arr, lo, h = cpu.scratch(b)
        g = eb_geometry(rank, lo, h, dx, origin, phi, nsub=nsub)
        t = eb_fab_type(g)
        cpu.set_eb(b, g, t)
        x = origin[0] + (np.arange(lo[0], h[0] + 1) + 0.5) * dx[0]
        shape = arr.shape[1:]
        rho = rho0 + np.zeros(shape)
        p = np.where(x < shock_x, p_ratio * 101325.0, 101325.0)[None, None, :] + np.zeros(shape)
        st = complete_from_prim(rank, rho, [np.zeros(shape)] * rank, p) * (g["alpha"][0] > 0)
        arr[...] = st
        if gpu is not None:
            gpu.set_eb(b, g, t)
            gpu.upload_state(b, st)
        geoms.append(g)
    cpu.fill_ghost_layer()
    if gpu is not None:
        gpu.fill_ghost_layer()
    return cpu, gpu, geoms


# BASELINE configs 4 and 5 in small: the geometries of the reference's examples
NOZZLE_CASE = dict(rank=2, n=(96, 48), origin=[-0.12, -0.0612, 0.0], extent=(0.24, 0.12),
                   method=(0, 1, 2), gcw=(8, 6), shock_x=-0.08, p_ratio=5.0,
                   boxes=[([0, 0, 0], [47, 47, 0]), ([48, 0, 0], [95, 47, 0])])
# (the whole cross-section: the y / z boundaries lie in the covered region outside the outer
# cylinder and the x boundaries cut walls that are parallel to x, so the transmissive ghost
# cells agree with the ghost geometry)
PLENUM_CASE = dict(rank=3, n=(48, 48, 48), origin=[-0.1, -0.2, -0.2], extent=(0.4, 0.4, 0.4),
                   method=(0, 2, 0), gcw=(4, 2), shock_x=-0.05, p_ratio=3.0,
                   boxes=[([0, 0, 0], [23, 47, 47]), ([24, 0, 0], [47, 47, 47])])
