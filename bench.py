#!/usr/bin/env python
"""bench.py -- fp64 cell-updates/s of the split-sweep patch update.

Workload (BASELINE.json configs[1]): 3-D PerfectGas spherical
Riemann problem, gamma = 1.4, MUSCL-Hancock<HLLEM(Larrouturou), VanLeer>,
Strang splitting (X Y Z Z Y X, half steps), scratch/flux ghost widths 8/6,
transmissive boundaries, CFL 0.5.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S]
                    [--scaling weak|strong] [--impl reference]

* weak scaling (default): one 512^3 box per GPU, the domain grows with N;
* strong scaling: the 512^3 domain is cut into N boxes (8 x 256^3 on 8 GPUs).  Every
  weak run with N > 1 also times the strong decomposition and reports it under
  "strong" in the same JSON line.

A "step" is one coarse time step: the all-direction stable-dt pass, six fused
sweeps (each = stable dt + numeric flux + conservative update + complete) and
the ghost refill; it advances every valid cell by 6 cell-updates.

Before the timed region every run checks parity: one coarse step, then a 40^3
sub-box that straddles the rank boundaries (N > 1: it crosses the NCCL halo)
and the discontinuity is recomputed by the CPU oracle from the same initial
data and compared BIT FOR BIT with the device state ("parity_check"); a
mismatch aborts the run with a non-zero exit code.

--impl reference times the reference's CPU path (the oracle restatement; the
upstream binary cannot be built here) on all host cores of rank 0,
on a bounded sample of the same workload: one 256^3 box.

Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fp64 cell-updates/s (3D PerfectGas MUSCL+HLLEM)"
UNIT = "cell-updates/s"
BYTES_PER_CELL_UPDATE = 112  # 2 * (3 + 4) * 8 B, SURVEY 8d
SWEEPS_PER_STEP = 6          # 3-D Strang
GCW = (8, 6)
CFL = 0.5
# box edge of the CPU sample; the environment variable is for smoke tests only
REFERENCE_SIZE = int(os.environ.get("FVS_REFERENCE_SIZE", "256"))


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def grid_parts(n_gpus):
    parts = [1, 1, 1]
    d = 0
    n = n_gpus
    while n > 1:
        parts[d % 3] *= 2
        n //= 2
        d += 1
    return parts


def grid_of_boxes(n_gpus, size, strong=False):
    """Weak: n_gpus boxes of size^3 arranged 2x2x2-style.  Strong: the size^3 domain cut
    into n_gpus boxes.  Returns (domain_n, boxes, parts, box_n)."""
    parts = grid_parts(n_gpus)
    box_n = [size // parts[d] for d in range(3)] if strong else [size] * 3
    boxes = []
    for k in range(parts[2]):
        for j in range(parts[1]):
            for i in range(parts[0]):
                lo = [i * box_n[0], j * box_n[1], k * box_n[2]]
                boxes.append((lo, [lo[d] + box_n[d] - 1 for d in range(3)]))
    return [parts[d] * box_n[d] for d in range(3)], boxes, parts, box_n


def geometry(domain_n, size):
    """The size^3 problem spans [-1, 1]^3 (dx = 2 / size); the domain is centred at 0."""
    dx = [2.0 / size] * 3
    origin = [-0.5 * domain_n[d] * dx[d] for d in range(3)]
    return dx, origin


def lattice_state(lo, hi, dx, origin, period):
    """Spherical Riemann problem (r < 0.4: rho = p = 1, else 0.125 / 0.1, at rest) on the fab
    box [lo, hi].  period[d] = 0: one sphere at the centre of the domain (1 GPU, and the strong-
    scaling decomposition).  period[d] = 2 (weak scaling, directions cut into boxes): the sphere
    repeats with the box size, centred ON the box interfaces -- every box holds the same amount of
    sphere and of wave front as the 1-GPU problem (per-GPU work really is fixed; with one sphere
    scaled to the grown domain the lazy stable dt and the zero-slope fast paths see different data
    and the sweeps of a 2-GPU run are 2 % slower for that reason alone), and the front crosses
    the rank boundaries, which is where the parity check looks."""
    from helpers import cell_centers, complete_from_prim
    z, y, x = cell_centers(lo, hi, dx, origin)
    shape = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    r2 = np.zeros(shape)
    for c, per in zip((x, y, z), period):
        if per:
            c = np.mod(c + 0.5 * per, per) - 0.5 * per
        r2 = r2 + c * c
    inside = r2 < 0.4 * 0.4
    rho = np.where(inside, 1.0, 0.125)
    p = np.where(inside, 1.0, 0.1)
    zero = np.zeros(shape)
    return complete_from_prim(3, rho, [zero] * 3, p)


def box_state(lo, hi, dx, origin, period, out=None):
    """lattice_state built plane by plane to bound host memory."""
    shape = (7, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    if out is None:
        out = np.empty(shape)
    for k in range(lo[2], hi[2] + 1):
        out[:, k - lo[2]] = lattice_state([lo[0], lo[1], k], [hi[0], hi[1], k], dx, origin, period)[:, 0]
    return out


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU path ----------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def oracle_level(orc, n):
    from helpers import riemann_state
    dx = [2.0 / n] * 3
    lvl = orc.Level(3, [0] * 3, [n - 1] * 3, dx, [([0] * 3, [n - 1] * 3)],
                    scratch_gcw=GCW[0], flux_gcw=GCW[1])
    lvl.set_method(orc.BASE_HLLEM_LARROUTUROU, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_VANLEER)
    for d in range(3):
        for s in (0, 1):
            lvl.add_bc(d, s, orc.BC_TRANSMISSIVE)
    arr, lo, hi = lvl.scratch(0)
    arr[...] = riemann_state(3, lo, hi, dx, [-1.0] * 3)
    lvl.fill_ghost_layer()
    return lvl


def cpu_reference_run(steps, warmup, size=REFERENCE_SIZE, budget_s=30.0):
    """The reference's CPU path (oracle port: 4 passes per sweep, OpenMP over tiles, 8-wide
    chunks) on all host cores, on a bounded sample of the workload: one size^3 box.

    Three builds of the same source exist: the strict one the parity tests use (AVX2, no
    FMA contraction) and two timing builds with the compiler's default contraction (AVX2 +
    FMA, AVX-512) -- what a Release build of the reference would use.  Every build the host
    CPU supports is timed on one step; the fastest one is the baseline."""
    from oracle import oracle as orc
    orc.build()
    cores = host_cores()
    flags = cpu_flags()
    builds = {"strict (x86-64-v3, -ffp-contract=off)": None}
    timing = orc.build_timing()
    if "v3fma" in timing and {"avx2", "fma"} <= flags:
        builds["x86-64-v3 + FMA contraction"] = timing["v3fma"]
    if "v4" in timing and {"avx512f", "avx512dq", "avx512vl", "avx512bw", "avx512cd"} <= flags:
        builds["x86-64-v4 (AVX-512) + FMA contraction"] = timing["v4"]
    rates = {}
    for name, path in builds.items():
        orc.unload()
        if path:
            orc.use_library(path)
        orc.set_num_threads(cores)
        lvl = oracle_level(orc, size)
        lvl.run(CFL, 1e9, 1, orc.SPLIT_STRANG)                 # touch the pages
        t0 = time.perf_counter()
        lvl.run(CFL, 1e9, 2, orc.SPLIT_STRANG)
        rates[name] = SWEEPS_PER_STEP * size ** 3 / (time.perf_counter() - t0)
        del lvl
    best = max(rates, key=rates.get)
    orc.unload()
    if builds[best]:
        orc.use_library(builds[best])
    threads = orc.set_num_threads(cores)
    lvl = oracle_level(orc, size)
    cycles = 0
    for _ in range(max(warmup, 1)):
        cycles += 1
        lvl.run(CFL, 1e9, cycles, orc.SPLIT_STRANG)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cycles += 1
        lvl.run(CFL, 1e9, cycles, orc.SPLIT_STRANG)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    del lvl
    orc.unload()                                                  # parity checks use the strict build
    updates = done * SWEEPS_PER_STEP * size ** 3
    grown = size + 2 * GCW[0]
    return {
        "value": updates / el, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{done} Strang steps of the same problem on one {size}^3 box (gcw {GCW[0]}/{GCW[1]}: "
                  f"the sweeps cover {grown}^3 = {(grown / size) ** 3:.2f}x the valid cells, 1.10x at 512^3; "
                  f"valid cell-updates counted), oracle restatement of the reference's CPU path, "
                  f"{threads} OpenMP threads, {el:.1f} s",
        "build": best,
        "builds_timed": {k: round(v) for k, v in rates.items()},
    }, el, done


# ---- device side -----------------------------------------------------------------
class Job:
    """One decomposition of the workload on the ranks of this run."""

    def __init__(self, env, domain_n, boxes, parts, box_n, arithmetic, size, strong=False):
        capi, torch, dist = env["capi"], env["torch"], env["dist"]
        rank, world, local_rank = env["rank"], env["world"], env["local_rank"]
        self.env, self.domain_n, self.boxes, self.parts, self.box_n = env, domain_n, boxes, parts, box_n
        self.dx, self.origin = geometry(domain_n, size)
        # weak scaling: the sphere repeats with the box size along the cut directions
        self.period = [0.0 if strong or parts[d] == 1 else 2.0 for d in range(3)]
        self.owner = list(range(len(boxes))) if world > 1 else [0] * len(boxes)
        self.arithmetic = arithmetic
        self.ctx = self.new_context()
        self.my_boxes = self.ctx.owned_boxes()
        self.stream = torch.cuda.ExternalStream(self.ctx.stream)

    def new_context(self):
        env = self.env
        capi, torch, dist = env["capi"], env["torch"], env["dist"]
        rank, world = env["rank"], env["world"]
        c = capi.Context(3, [0] * 3, [d - 1 for d in self.domain_n], self.dx, self.boxes,
                         scratch_gcw=GCW[0], flux_gcw=GCW[1], device=env["local_rank"],
                         comm_rank=rank if world > 1 else 0, comm_size=world, owner=self.owner)
        if world > 1:
            idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            c.comm_init(bytes(idt.cpu().numpy().tobytes()))
        for d in range(3):
            for s in (0, 1):
                c.add_bc(d, s, capi.BC_TRANSMISSIVE)
        c.set_arithmetic(self.arithmetic)
        return c

    def fill(self, keep_host=False):
        """Uploads the initial state; returns the pinned host copies if asked to keep them."""
        torch = self.env["torch"]
        host = {}
        for b in self.my_boxes:
            lo, hi, _ = self.ctx.state_box(b)
            t = torch.empty(self.ctx.state_shape(b), dtype=torch.float64, pin_memory=True)
            box_state(lo, hi, self.dx, self.origin, self.period, out=t.numpy())
            self.ctx.upload_state(b, t.numpy())
            if keep_host:
                host[b] = t
        self.ctx.fill_ghost_layer()
        return host

    def barrier(self):
        torch, dist = self.env["torch"], self.env["dist"]
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        torch, dist = self.env["torch"], self.env["dist"]
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_step(self, ctx=None):
        capi = self.env["capi"]
        c = ctx or self.ctx
        dt = CFL * c.compute_stable_dt_all()
        ok, fail, _ = c.advance_hierarchy(dt, capi.SPLIT_STRANG)
        if not ok:  # CFL retry, RunSimulation.hpp:141-151
            ok, fail, _ = c.advance_hierarchy(CFL * fail, capi.SPLIT_STRANG)
            assert ok
        return dt

    def timed(self, steps):
        """steps coarse steps between barriers; CUDA events on the context's stream, max over ranks."""
        torch = self.env["torch"]
        self.ctx.enable_timing(True)
        launches0 = self.ctx.kernel_launches
        s0, n0 = self.ctx.sweep_time_ms()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record(self.stream)
        for _ in range(steps):
            self.one_step()
        ev1.record(self.stream)
        self.barrier()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1))
        s1, n1 = self.ctx.sweep_time_ms()
        self.ctx.enable_timing(False)
        return ms, self.ctx.kernel_launches - launches0, s1 - s0, n1 - n0

    def total_cells(self):
        cells = self.box_n[0] * self.box_n[1] * self.box_n[2]
        return cells * (len(self.boxes) if self.env["world"] > 1 else len(self.my_boxes))

    def parity_check(self):
        """One coarse step on the device; a 40^3 sub-box across the rank boundaries and the
        discontinuity recomputed by the oracle from the same initial data must equal the
        device state bit for bit (the wide-halo cadence makes a box independent of its
        neighbours within one step: tests/test_gpu_fullsize.py)."""
        from oracle import oracle as orc
        torch, dist = self.env["torch"], self.env["dist"]
        orc.build()
        orc.set_num_threads(min(host_cores(), 8))
        dn, parts = self.domain_n, self.parts
        centre = [dn[d] // 2 for d in range(3)]
        radius = int(round(0.4 / self.dx[0]))
        # on every internal boundary plane where there is one, on the sphere along another axis
        t = [self.box_n[d] if parts[d] > 1 else centre[d] for d in range(3)]
        free = [d for d in range(3) if parts[d] == 1]
        axis = free[-1] if free else 2
        t[axis] = centre[axis] + radius
        sub_lo = [max(0, min(t[d] - 20, dn[d] - 40)) for d in range(3)]
        sub_hi = [sub_lo[d] + 39 for d in range(3)]
        dt = self.one_step()
        lvl = orc.Level(3, [0] * 3, [d - 1 for d in dn], self.dx, [(sub_lo, sub_hi)],
                        scratch_gcw=GCW[0], flux_gcw=GCW[1])
        lvl.set_method(orc.BASE_HLLEM_LARROUTUROU, orc.RECON_MUSCL_HANCOCK, orc.LIMITER_VANLEER)
        arr, slo, shi = lvl.scratch(0)
        arr[...] = lattice_state(slo, shi, self.dx, self.origin, self.period)
        ok, _, _ = lvl.advance_hierarchy(dt, orc.SPLIT_STRANG)
        g = GCW[0]
        ref = arr[(slice(None),) + (slice(g, -g),) * 3]
        bad = 0 if ok else 1
        cells = 0
        moved = 0
        for b in self.my_boxes:
            blo, bhi = self.boxes[b]
            lo = [max(sub_lo[d], blo[d]) for d in range(3)]
            hi = [min(sub_hi[d], bhi[d]) for d in range(3)]
            if any(lo[d] > hi[d] for d in range(3)):
                continue
            got = self.ctx.download_state_region(b, lo, hi)
            sl = (slice(None),) + tuple(slice(lo[d] - sub_lo[d], hi[d] - sub_lo[d] + 1) for d in (2, 1, 0))
            want = ref[sl]
            bad += int((got.view(np.uint64) != np.ascontiguousarray(want).view(np.uint64)).sum())
            cells += got[0].size
            moved += int((want[1:4] != 0).any())
        if dist is not None:
            v = torch.tensor([bad, cells, moved], dtype=torch.int64, device="cuda")
            dist.all_reduce(v)
            bad, cells, moved = int(v[0]), int(v[1]), int(v[2])
        ranks = sum(1 for (blo, bhi) in self.boxes
                    if all(max(sub_lo[d], blo[d]) <= min(sub_hi[d], bhi[d]) for d in range(3)))
        return {"subbox": [sub_lo, sub_hi], "bitwise": bad == 0, "values_differing": bad,
                "cells_compared": cells, "boxes_touched": ranks, "sees_the_wave": moved > 0,
                "dt": dt, "oracle": "strict build, same initial data, one Strang step"}

    def close(self):
        self.ctx.close()


def run_e2e(job, data_in, e2e_steps, pipelined):
    """End to end through the C ABI with host buffers.  Every e2e step is one independent
    batch: the pinned level data of the rank's boxes (valid cells,
    IntegratorContext::CopyDataToScratch) go to the device, ghost fill + one coarse step
    run there, the advanced level data (CopyScratchToData) come back to a pinned output
    buffer.  Pipelined: three contexts (upload of batch k+1, compute of batch k and
    download of batch k-1 overlap: PCIe is full duplex) over two output buffers."""
    env = job.env
    torch = env["torch"]
    my_boxes = job.my_boxes
    n_pipe = 3 if pipelined else 1
    pipe = [job.ctx] + [job.new_context() for _ in range(n_pipe - 1)]
    streams = [torch.cuda.ExternalStream(c.stream) for c in pipe]
    if n_pipe > 1:
        outs = [{b: torch.empty_like(data_in[b]).pin_memory() for b in my_boxes} for _ in range(2)]
    else:
        outs = [data_in]
    trace = os.environ.get("FVS_E2E_TRACE")                  # host time line of every batch (debug)

    def upload(c):
        for b in my_boxes:
            c.copy_data_to_scratch_async(b, data_in[b].numpy())   # H2D from pinned memory

    def run(nsteps):
        done = [None, None]                                   # download events per output buffer
        upload(pipe[0])
        t_prev = time.perf_counter()
        for k in range(nsteps):
            cur = pipe[k % n_pipe]
            if k + 1 < nsteps and n_pipe > 1:
                upload(pipe[(k + 1) % n_pipe])               # overlaps cur's compute
            if trace:
                cur.synchronize()
                t_up = time.perf_counter()
            cur.fill_ghost_layer()
            job.one_step(cur)                                # synchronises cur's stream (dt)
            if trace and env["rank"] == 0:
                t_now = time.perf_counter()
                print(f"e2e batch {k}: waited for upload {1e3 * (t_up - t_prev):.1f} ms, "
                      f"compute {1e3 * (t_now - t_up):.1f} ms", file=sys.stderr)
                t_prev = t_now
            o = k % len(outs)
            if done[o % 2] is not None:
                done[o % 2].synchronize()                    # the buffer's previous batch has landed
            for b in my_boxes:
                cur.copy_scratch_to_data_async(b, outs[o][b].numpy())   # D2H of the result
            if n_pipe > 1:
                ev = torch.cuda.Event()
                ev.record(streams[k % n_pipe])
                done[o % 2] = ev
            if n_pipe == 1 and k + 1 < nsteps:
                cur.synchronize()
                upload(cur)
        for c in pipe:
            c.synchronize()

    run(n_pipe)                                              # every context once: buffers allocated
    job.barrier()
    t0 = time.perf_counter()
    run(e2e_steps)
    job.barrier()
    e2e_ms = job.max_over_ranks(1e3 * (time.perf_counter() - t0))
    for c in pipe[1:]:
        c.close()
    return e2e_ms, n_pipe


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512,
                    help="weak: cells per box edge (one box per GPU); strong: edge of the whole domain")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=24,
                    help="independent batches timed end to end; the three-stage pipeline needs "
                         "about two batches to fill and drain, so few batches understate its rate")
    ap.add_argument("--no-e2e-pipeline", dest="e2e_pipeline", action="store_false",
                    help="one context, upload / step / download strictly in sequence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the AMR (W3) and cut-cell (W4) sub-records of the N=1 line")
    ap.add_argument("--arithmetic", default="ieee", choices=["ieee", "tolerance"],
                    help="arithmetic of the fused sweep for the headline numbers: ieee = "
                         "bit-identical to the oracle (default); tolerance = within 1e-12 per step "
                         "(fvs_set_arithmetic).  The other one is timed too and reported under "
                         "'other_arithmetic'")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, 1)
    strong = args.scaling == "strong"

    domain_n, boxes, parts, box_n = grid_of_boxes(n_gpus, args.size, strong)
    workload = (f"3D spherical Riemann problem, PerfectGas gamma=1.4, "
                f"MUSCL-Hancock<HLLEM-Larrouturou,VanLeer>, Strang split sweeps, "
                f"{domain_n[0]}x{domain_n[1]}x{domain_n[2]} cells = {n_gpus} box(es) of "
                f"{box_n[0]}x{box_n[1]}x{box_n[2]}"
                + (", the sphere repeated with the box size and centred on the box interfaces"
                   if n_gpus > 1 and not strong else "")
                + f", gcw {GCW[0]}/{GCW[1]}, transmissive, CFL {CFL}")
    config = {
        "workload": workload,
        "cells_per_gpu": box_n[0] * box_n[1] * box_n[2], "sweeps_per_step": SWEEPS_PER_STEP,
        "l2_policy": "inputs larger than L2 (8.2 GB state per sweep at 512^3)",
        "partition": f"{n_gpus} box(es), one per GPU",
    }

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, el, done = cpu_reference_run(args.steps, min(args.warmup, 1), budget_s=120.0)
        n = REFERENCE_SIZE
        config = {
            "workload": f"3D spherical Riemann problem, PerfectGas gamma=1.4, "
                        f"MUSCL-Hancock<HLLEM-Larrouturou,VanLeer>, Strang split sweeps, "
                        f"{n}x{n}x{n} cells = 1 box of {n}^3, gcw {GCW[0]}/{GCW[1]}, transmissive, CFL {CFL}",
            "bounded_sample_of": workload,
            "cells_in_sample": n ** 3, "sweeps_per_step": SWEEPS_PER_STEP,
            "host": f"rank 0, {base['cores']} OpenMP threads, build: {base['build']}",
        }
        line = {
            "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": n_gpus, "steps": done, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * el / max(done, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    from finitevolumesolver_b200 import capi

    if capi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    if world != n_gpus:
        raise SystemExit(f"--gpus {n_gpus} needs {n_gpus} ranks (torchrun --nproc-per-node {n_gpus}); "
                         f"this run has WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = {"capi": capi, "torch": torch, "dist": dist, "rank": rank, "world": world,
           "local_rank": local_rank}

    arith = {"ieee": capi.ARITH_IEEE, "tolerance": capi.ARITH_TOLERANCE}
    other_name = "tolerance" if args.arithmetic == "ieee" else "ieee"
    config["arithmetic"] = (
        "ieee: every fp64 operation correctly rounded, no FMA contraction, bit-identical to the "
        "CPU oracle" if args.arithmetic == "ieee" else
        "tolerance: reciprocal-multiply quotients + FMA contraction, within 1e-12 relative per "
        "step of the CPU oracle (tests/test_gpu_tolerance_arithmetic.py)")

    job = Job(env, domain_n, boxes, parts, box_n, arith[args.arithmetic], args.size, strong)
    host = job.fill(keep_host=True)
    my_boxes = job.my_boxes
    ctx = job.ctx

    # ---- parity before anything is timed (always in the bit-identical arithmetic) -----
    parity = None
    if not args.no_parity_check and min(box_n) >= 64:
        ctx.set_arithmetic(capi.ARITH_IEEE)
        parity = job.parity_check()
        ctx.set_arithmetic(arith[args.arithmetic])
        if not parity["bitwise"]:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "error": "device state differs from the oracle"}))
            return 3

    for _ in range(max(args.warmup, 3)):
        job.one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    redos0 = ctx.dt_hint_redos
    elapsed_ms, launches, sweep_ms, n_sweeps = job.timed(args.steps)
    redos_timed = ctx.dt_hint_redos - redos0
    clocks = sampler.stop() if rank == 0 else None
    total_cells = job.total_cells()
    updates = args.steps * SWEEPS_PER_STEP * total_cells
    value = updates / (elapsed_ms * 1e-3)

    # ---- the same steps with the other arithmetic (secondary, not the headline) ----
    ctx.set_arithmetic(arith[other_name])
    for _ in range(3):
        job.one_step()
    other_ms, _, o_sweep_ms, o_n = job.timed(args.steps)
    other_sweep_s = o_sweep_ms / max(o_n, 1) * 1e-3
    ctx.set_arithmetic(arith[args.arithmetic])

    # ---- end to end ------------------------------------------------------------------
    e2e_steps = max(args.e2e_steps, 1)
    g = GCW[0]
    data_in = {b: host[b][:, g:-g, g:-g, g:-g].contiguous().pin_memory() for b in my_boxes}
    host.clear()
    bytes_state = sum(data_in[b].numel() * 8 for b in my_boxes)
    # the pipeline pins 3 x the level data per rank (input + two output buffers); fall back
    # to the sequential form if the host cannot give that to every rank of this node
    pipelined = args.e2e_pipeline
    if pipelined:
        try:
            import psutil
            need = 2.2 * bytes_state * world
            pipelined = psutil.virtual_memory().available > need
        except Exception:
            pipelined = world <= 2
    if dist is not None:
        t = torch.tensor([1 if pipelined else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        pipelined = bool(int(t.item()))
    e2e_ms, n_pipe = run_e2e(job, data_in, e2e_steps, pipelined)
    e2e_value = e2e_steps * SWEEPS_PER_STEP * total_cells / (e2e_ms * 1e-3)
    data_in.clear()

    # ---- roofline of the dominant kernel (the fused sweep) --------------------
    peak, peak_src = measured_peak_gbs()
    per_launch_bytes = BYTES_PER_CELL_UPDATE * len(my_boxes) * box_n[0] * box_n[1] * box_n[2]
    avg_sweep_s = (sweep_ms / max(n_sweeps, 1)) * 1e-3
    achieved = per_launch_bytes / avg_sweep_s / 1e9 if n_sweeps else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("size") == box_n[0] == box_n[1] == box_n[2] and len(my_boxes) == 1:
                traffic = tj.get("fused_sweep_bytes_per_launch" if args.arithmetic == "ieee"
                                 else "fused_sweep_bytes_per_launch_tolerance")
                traffic_src = ("profiles/traffic.json: ncu --set full capture of the same command, "
                               "not measured in this run")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "kernel": ("fvs::" if args.arithmetic == "ieee" else "fvs_tol::") +
                  "fused_sweep_x_smem_kernel / fused_sweep_yz_smem_kernel<3,*,HLLEM-L,MUSCL,VanLeer>",
        "avg_launch_ms": 1e3 * avg_sweep_s, "launches_timed": n_sweeps,
        "share_of_step": sweep_ms / elapsed_ms if elapsed_ms else None,
        "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
    }
    other = {
        "arithmetic": other_name, "value": updates / (other_ms * 1e-3), "unit": UNIT,
        "ms_per_step": other_ms / args.steps, "avg_launch_ms": 1e3 * other_sweep_s,
        "roofline_frac": (per_launch_bytes / other_sweep_s / 1e9 / peak) if other_sweep_s else None,
        "note": "same workload and steps, device-resident, timed right after the headline region",
    }
    hint_redos = ctx.dt_hint_redos
    job.close()

    # ---- strong scaling of the same 512^3 domain (sub-record of every weak run) -------
    strong_rec = None
    if not strong and n_gpus > 1 and not args.no_strong:
        sdom, sboxes, sparts, sbox_n = grid_of_boxes(n_gpus, args.size, True)
        sjob = Job(env, sdom, sboxes, sparts, sbox_n, arith[args.arithmetic], args.size, True)
        sjob.fill()
        sparity = None
        if not args.no_parity_check and min(sbox_n) >= 64:
            sparity = sjob.parity_check()
            if not sparity["bitwise"]:
                if rank == 0:
                    print(json.dumps({"parity_check": sparity, "error": "strong decomposition differs"}))
                return 3
        for _ in range(max(args.warmup, 3)):
            sjob.one_step()
        s_ms, _, s_sweep_ms, s_n = sjob.timed(args.steps)
        grown = [sbox_n[d] + 2 * GCW[0] for d in range(3)]
        strong_rec = {
            "value": args.steps * SWEEPS_PER_STEP * sjob.total_cells() / (s_ms * 1e-3), "unit": UNIT,
            "ms_per_step": s_ms / args.steps, "avg_launch_ms": s_sweep_ms / max(s_n, 1),
            "workload": f"{sdom[0]}x{sdom[1]}x{sdom[2]} cells = {n_gpus} box(es) of "
                        f"{sbox_n[0]}x{sbox_n[1]}x{sbox_n[2]}",
            "ghost_overcompute": (grown[0] * grown[1] * grown[2]) / (sbox_n[0] * sbox_n[1] * sbox_n[2]),
            "parity_check": sparity,
            "note": "total work fixed: compare with the N=1 value of the same command; the sweeps "
                    "cover the grown box (wide-halo cadence), so the per-GPU work falls slower than 1/N",
        }
        sjob.close()

    # ---- secondary workloads (N = 1 only): AMR hierarchy (W3) and cut-cell level (W4) ----
    secondary = {}
    if n_gpus == 1 and not args.no_secondary:
        import bench_secondary
        for name, fn in (("amr", bench_secondary.amr_record), ("eb", bench_secondary.eb_record)):
            try:
                secondary[name] = fn(capi, torch, peak)
            except Exception as e:  # a sub-record must not take the headline down
                secondary[name] = {"error": f"{type(e).__name__}: {e}"}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    cpu_baseline = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_run(10, 1, budget_s=20.0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "clocks": clocks,
        # whole job: every rank moves its own boxes
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_state * max(world, 1),
                "d2h_bytes_per_step": bytes_state * max(world, 1), "steps": e2e_steps,
                "pipeline": f"{n_pipe} context(s) per rank; every step is an independent batch "
                            "(pinned level data = valid cells in, ghost fill + dt pass + 6 sweeps on the "
                            "device, level data out: CopyDataToScratch / CopyScratchToData); "
                            "wall clock around all batches"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "other_arithmetic": other, "parity_check": parity,
        # lazy stable dt: evaluations repeated because a hint was violated
        "dt_hint_redos": {"timed_region": redos_timed, "whole_run": hint_redos},
    }
    if strong_rec is not None:
        line["strong"] = strong_rec
    line.update(secondary)
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
