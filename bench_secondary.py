"""Secondary bench records (reported inside bench.py's JSON line, never the headline):

  amr  -- SURVEY 8d W3 / BASELINE config 3: 2-D diagonal Sod (examples/AMReX/SodShockTube2d.cpp:57-64)
          on three levels, ratio 2, GradientDetector(density, 1e-2) + TagBuffer, blocking factor 8,
          max grid 128, reflective boundaries, Strang splitting, fine-first subcycling with flux
          correction; the hierarchy comes from the device tags through fvs_cluster_tags /
          fvs_regrid_level.  Timed twice: coarse steps on the static hierarchy (CUDA events on the
          context's stream) and coarse steps that regrid first (wall clock: the tags go through
          the host).
  eb   -- SURVEY 8d W4: the Schardin wedge (examples/AMReX/EB/2D/Schardin.cpp:87-130) as a
          multi-box cut-cell level, KbnCutCellMethod(MusclHancock<HLLEM, VanLeer>, HLL-Einfeldt),
          gcw 8/6, Strang, CFL 0.4.

A cell-update is one cell advanced through one directional sweep; the roofline
figure is 2 (N + 4) 8 B = 96 B per 2-D cell-update (cut cells: + 8 (15 + 2N) B of geometry).
Nothing here touches oracle/ -- parity of these paths is tests/test_gpu_regrid.py,
tests/test_gpu_amr.py and tests/test_gpu_cutcell.py."""
import time

import numpy as np

GAMMA = 1.4


def complete_from_prim(rank, rho, vel, p):
    mom = [rho * v for v in vel]
    e = p / (GAMMA - 1.0) + 0.5 * rho * sum(v * v for v in vel)
    a = np.sqrt(GAMMA * p / rho)
    return np.stack([rho] + mom + [e, p, a])


# ---------------------------------------------------------------------------------- AMR (W3)
def _sod(lo, hi, dx):
    x = -1.0 + (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    y = -1.0 + (np.arange(lo[1], hi[1] + 1) + 0.5) * dx[1]
    upper = (x[None, None, :] + y[None, :, None]) > 0.0
    rho = np.where(upper, 0.125, 1.0)
    p = np.where(upper, 10132.5, 101325.0)
    return complete_from_prim(2, rho, [np.zeros_like(rho)] * 2, p)


class AmrRun:
    RANK, GCW, BF, MAXGRID, NEST, BUFFER, LEVELS = 2, (8, 6), 8, 128, 1, 4, 3

    def __init__(self, capi, n0):
        self.capi, self.n0 = capi, n0
        hi, dx = self.geom(0)
        half = n0 // 2
        boxes = [([i * half, j * half, 0], [(i + 1) * half - 1, (j + 1) * half - 1, 0])
                 for j in range(2) for i in range(2)]
        self.ctx = capi.Context(2, [0] * 3, hi, dx, boxes, scratch_gcw=self.GCW[0], flux_gcw=self.GCW[1])
        for d in range(2):
            for s in (0, 1):
                self.ctx.add_bc(d, s, capi.BC_REFLECTIVE)
        self.levels = [self.ctx]
        self.fill(self.ctx, 0)
        self.regrid(init=True)

    def geom(self, l):
        f = 2 ** l
        return [self.n0 * f - 1, self.n0 * f - 1, 0], [2.0 / (self.n0 * f), 2.0 / (self.n0 * f), 1.0]

    def fill(self, lv, l):
        _, dx = self.geom(l)
        for b in range(len(lv.boxes)):
            lo, hi, _ = lv.state_box(b)
            lv.upload_state(b, _sod(lo, hi, dx))
        lv.fill_ghost_layer()

    def tags(self, l):
        hi, _ = self.geom(l)
        return self.levels[l].tag_level([hi[0] + 1, hi[1] + 1, 1], [(0, 1e-2)],
                                        buffer_width=self.BUFFER)

    def regrid(self, init=False):
        l = 0
        while l + 1 < self.LEVELS and l < len(self.levels):
            hi, _ = self.geom(l)
            boxes = self.capi.cluster_tags(2, [0] * 3, hi, [0] * 3, self.levels[l].boxes, self.tags(l),
                                           blocking_factor=self.BF, max_grid_size=self.MAXGRID,
                                           nesting_buffer=self.NEST)
            if not boxes:
                self.ctx.regrid_level(l + 1, [])
                del self.levels[l + 1:]
                break
            g = self.ctx.regrid_level(l + 1, boxes)
            if init:
                self.fill(g, l + 1)
            if l + 1 < len(self.levels):
                self.levels[l + 1] = g
            else:
                self.levels.append(g)
            l += 1

    def step(self):
        dt = 0.5 * self.ctx.compute_stable_dt_hierarchy()
        ok, fail, _ = self.ctx.advance_hierarchy(dt, self.capi.SPLIT_STRANG)
        if not ok:
            ok, _, _ = self.ctx.advance_hierarchy(0.5 * fail, self.capi.SPLIT_STRANG)
            assert ok

    def updates_per_coarse_step(self):
        total = 0
        for l, lv in enumerate(self.levels):
            cells = sum((h[0] - lo[0] + 1) * (h[1] - lo[1] + 1) for lo, h in lv.boxes)
            total += cells * 4 * 2 ** l
        return total


def amr_record(capi, torch, peak_gbs, n0=512, steps=20):
    run = AmrRun(capi, n0)
    stream = torch.cuda.ExternalStream(run.ctx.stream)
    for _ in range(3):
        run.step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0 = run.ctx.kernel_launches
    upd = run.updates_per_coarse_step()
    ev0.record(stream)
    for _ in range(steps):
        run.step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = (run.ctx.kernel_launches - l0) / steps
    boxes = [len(lv.boxes) for lv in run.levels]
    cells = [sum((h[0] - lo[0] + 1) * (h[1] - lo[1] + 1) for lo, h in lv.boxes) for lv in run.levels]
    # the same with a regrid in front of every coarse step (W3: "regrid every coarse step")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    upd_r = 0
    for _ in range(steps):
        run.regrid()
        upd_r += run.updates_per_coarse_step()
        run.step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    value = steps * upd / (ms * 1e-3)
    rec = {
        "workload": f"2D diagonal Sod, base {n0}x{n0} in 4 boxes, {len(run.levels)} levels ratio 2 "
                    f"(boxes per level {boxes}, cells per level {cells}), MUSCL-Hancock<HLLEM-L,VanLeer>, "
                    f"Strang, fine-first subcycling with flux correction, gcw 8/6, reflective, CFL 0.5",
        "value": value, "unit": "cell-updates/s", "ms_per_coarse_step": ms / steps,
        "cell_updates_per_coarse_step": upd, "launches_per_coarse_step": launches,
        "host_syncs_per_coarse_step": "2 (hierarchy stable dt, end-of-step dt checks)",
        "roofline_frac_whole_step": value * 96.0 / 1e9 / peak_gbs,
        "with_regrid_every_step": {"value": upd_r / wall, "unit": "cell-updates/s",
                                   "ms_per_coarse_step": 1e3 * wall / steps,
                                   "note": "wall clock; tags from the device kernels, clustering on the host "
                                           "(fvs_cluster_tags), fvs_regrid_level"},
        "steps": steps,
    }
    run.ctx.close()
    return rec


# ----------------------------------------------------------------------------------- EB (W4)
WEDGE = [(0.02, 0.05), (0.05, 0.0655), (0.05, 0.0345)]


def polygon_phi(points):
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
        if (cx - x0) * nx + (cy - y0) * ny > 0:
            nx, ny = -nx, -ny
        edges.append((x0, y0, nx, ny))

    def phi(x, y):
        v = None
        for (x0, y0, nx, ny) in edges:
            t = (x - x0) * nx + (y - y0) * ny
            v = t if v is None else np.maximum(v, t)
        return v
    return phi


def eb_geometry_2d(lo, hi, dx, phi, nsub=8):
    """alpha, beta, centroid and normal arrays of one fab box from the level set (> 0 = fluid) by
    sub-sampling (stand-in for AMReX EB2; same construction as tests/helpers.py:eb_geometry).
    Boxes the interface does not touch are classified from the cell-centre values alone."""
    nx, ny = hi[0] - lo[0] + 1, hi[1] - lo[1] + 1
    xc = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
    yc = (np.arange(lo[1], hi[1] + 1) + 0.5) * dx[1]
    p0 = phi(xc[None, :], yc[:, None]) + np.zeros((ny, nx))
    far = 1.5 * max(dx[0], dx[1])
    if np.all(np.abs(p0) > far):
        fluid = float(p0.flat[0] > 0)
        if np.all((p0 > 0) == (fluid > 0)):
            alpha = np.full((1, 1, ny, nx), fluid)
            beta = [np.full((1, 1, ny, nx + 1), fluid), np.full((1, 1, ny + 1, nx), fluid)]
            z = np.zeros((2, 1, ny, nx))
            return {"alpha": alpha, "beta": beta, "bcent": z, "bnorm": z.copy()}, (0 if fluid else -1)
    off = (np.arange(nsub) + 0.5) / nsub
    xs = ((np.arange(lo[0], hi[0] + 1))[:, None] + off[None, :]) * dx[0]   # [nx, s]
    ys = ((np.arange(lo[1], hi[1] + 1))[:, None] + off[None, :]) * dx[1]
    xn = np.arange(lo[0], hi[0] + 2) * dx[0]
    yn = np.arange(lo[1], hi[1] + 2) * dx[1]
    alpha = (phi(xs[None, :, None, :], ys[:, None, :, None]) > 0).reshape(ny, nx, -1).mean(axis=2)
    bx = (phi(xn[None, :, None], ys[:, None, :]) > 0).mean(axis=2)            # [ny, nx+1]
    by = (phi(xs[None, :, :], yn[:, None, None]) > 0).mean(axis=2)            # [ny+1, nx]
    covered = alpha == 0.0
    bx[:, 1:] = np.where(covered, 0.0, bx[:, 1:])
    bx[:, :-1] = np.where(covered, 0.0, bx[:, :-1])
    by[1:, :] = np.where(covered, 0.0, by[1:, :])
    by[:-1, :] = np.where(covered, 0.0, by[:-1, :])
    h = 1e-4 * dx[0]
    X, Y = xc[None, :], yc[:, None]
    gx = (phi(X + h, Y) - phi(X - h, Y)) / (2 * h) + np.zeros_like(p0)
    gy = (phi(X, Y + h) - phi(X, Y - h)) / (2 * h) + np.zeros_like(p0)
    norm = np.sqrt(gx * gx + gy * gy)
    norm = np.where(norm > 0, norm, 1.0)
    cut = (alpha > 0) & (alpha < 1)
    bnorm = np.stack([np.where(cut, -gx / norm, 0.0), np.where(cut, -gy / norm, 0.0)])[:, None]
    bcent = np.stack([np.where(cut, np.clip(-p0 * (gx / norm) / norm / dx[0], -0.5, 0.5), 0.0),
                      np.where(cut, np.clip(-p0 * (gy / norm) / norm / dx[1], -0.5, 0.5), 0.0)])[:, None]
    geom = {"alpha": alpha[None, None], "beta": [bx[None, None], by[None, None]], "bcent": bcent,
            "bnorm": bnorm}
    if np.all(alpha == 0.0):
        return geom, -1
    if np.all(alpha == 1.0) and np.all(bx == 1.0) and np.all(by == 1.0):
        return geom, 0
    return geom, 1


def eb_record(capi, torch, peak_gbs, n=(2048, 1440), box=256, steps=10):
    extent = (0.15, 0.10)
    dx = [extent[0] / n[0], extent[1] / n[1], 1.0]
    boxes = []
    for j0 in range(0, n[1], box):
        for i0 in range(0, n[0], box):
            boxes.append(([i0, j0, 0], [min(i0 + box, n[0]) - 1, min(j0 + box, n[1]) - 1, 0]))
    ctx = capi.Context(2, [0] * 3, [n[0] - 1, n[1] - 1, 0], dx, boxes, scratch_gcw=8, flux_gcw=6)
    for d in range(2):
        for s in (0, 1):
            ctx.add_bc(d, s, capi.BC_TRANSMISSIVE)
    ctx.enable_eb()
    phi = polygon_phi(WEDGE)
    types = []
    cut_cells = 0
    fluid_cells = 0
    for b in range(len(boxes)):
        lo, hi, _ = ctx.state_box(b)
        geom, t = eb_geometry_2d(lo, hi, dx, phi)
        ctx.set_eb(b, geom, t)
        types.append(t)
        a = geom["alpha"][0, 0, 8:-8, 8:-8]
        cut_cells += int(((a > 0) & (a < 1)).sum())
        fluid_cells += int((a > 0).sum())
        x = (np.arange(lo[0], hi[0] + 1) + 0.5) * dx[0]
        shape = geom["alpha"].shape[1:]
        rho = np.ones(shape)
        p = np.where(x < 0.015, 5.0 * 101325.0, 101325.0)[None, None, :] + np.zeros(shape)
        st = complete_from_prim(2, rho, [np.zeros(shape)] * 2, p) * (geom["alpha"][0] > 0)
        ctx.upload_state(b, st)
    ctx.fill_ghost_layer()

    def step():
        dt = 0.4 * ctx.compute_stable_dt_all()
        ok, fail, _ = ctx.advance_hierarchy(dt, capi.SPLIT_STRANG)
        if not ok:
            ok, _, _ = ctx.advance_hierarchy(0.4 * fail, capi.SPLIT_STRANG)
            assert ok

    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    l0 = ctx.kernel_launches
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    cells = n[0] * n[1]
    value = steps * 4 * cells / (ms * 1e-3)
    bytes_per_step = 4 * (96.0 * cells + 8.0 * (15 + 4) * cut_cells)
    rec = {
        "workload": f"2D Schardin wedge, {n[0]}x{n[1]} cells in {len(boxes)} boxes of {box}^2 "
                    f"({types.count(1)} cut fabs, {types.count(0)} regular, {types.count(-1)} covered; "
                    f"{cut_cells} cut cells, {fluid_cells} fluid cells), KbnCutCellMethod(MUSCL-Hancock<HLLEM-L,"
                    f"VanLeer>, HLL-Einfeldt), Strang, gcw 8/6, CFL 0.4",
        "value": value, "unit": "cell-updates/s", "ms_per_step": ms / steps,
        "launches_per_step": (ctx.kernel_launches - l0) / steps,
        "roofline_frac_whole_step": bytes_per_step * steps / (ms * 1e-3) / 1e9 / peak_gbs,
        "algorithmic_bytes_per_step": bytes_per_step, "steps": steps,
    }
    ctx.close()
    return rec
