#!/usr/bin/env python
"""bench.py -- fp64 cell-updates/s of the split-sweep patch update.

Workload (BASELINE.json configs[1]): 3-D PerfectGas spherical
Riemann problem, gamma = 1.4, MUSCL-Hancock<HLLEM(Larrouturou), VanLeer>,
Strang splitting (X Y Z Z Y X, half steps), scratch/flux ghost widths 8/6,
transmissive boundaries, CFL 0.5; one 512^3 box per GPU (weak scaling: the
domain grows with the number of GPUs).

A "step" is one coarse time step: the all-direction stable-dt pass, six fused
sweeps (each = stable dt + numeric flux + conservative update + complete) and
the ghost refill; it advances every valid cell by 6 cell-updates.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

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


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def grid_of_boxes(n_gpus, size):
    """n_gpus boxes of size^3 arranged 2x2x2-style; returns (domain_n, boxes)."""
    parts = [1, 1, 1]
    d = 0
    n = n_gpus
    while n > 1:
        parts[d % 3] *= 2
        n //= 2
        d += 1
    boxes = []
    for k in range(parts[2]):
        for j in range(parts[1]):
            for i in range(parts[0]):
                lo = [i * size, j * size, k * size]
                boxes.append((lo, [lo[0] + size - 1, lo[1] + size - 1, lo[2] + size - 1]))
    return [parts[d] * size for d in range(3)], boxes


def box_state(lo, hi, dx, origin, out=None):
    """Spherical Riemann problem (r < 0.4: rho = p = 1, else 0.125 / 0.1, at
    rest) on the fab box [lo, hi]; built plane by plane to bound host memory."""
    from helpers import riemann_state
    shape = (7, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    if out is None:
        out = np.empty(shape)
    for k in range(lo[2], hi[2] + 1):
        out[:, k - lo[2]] = riemann_state(3, [lo[0], lo[1], k], [hi[0], hi[1], k], dx, origin)[:, 0]
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


def cpu_reference_run(steps, warmup, size=None, budget_s=20.0):
    """The reference's CPU path (oracle port, OpenMP over tiles + 8-wide
    chunks) on the host cores: bounded sample of the same workload."""
    from oracle import oracle as orc
    from helpers import riemann_state
    orc.build()
    n = size or 128
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
    cycles = 0
    for _ in range(warmup):
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
    updates = done * SWEEPS_PER_STEP * n ** 3
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {
        "value": updates / el, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{done} Strang steps of the same problem on one {n}^3 box "
                  f"(gcw {GCW[0]}/{GCW[1]}), oracle restatement of the reference's CPU path, "
                  f"{el:.1f} s",
    }, el, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512, help="cells per box edge (one box per GPU)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=24,
                    help="independent batches timed end to end; the three-stage pipeline needs "
                         "about two batches to fill and drain, so few batches understate its rate")
    ap.add_argument("--no-e2e-pipeline", dest="e2e_pipeline", action="store_false",
                    help="one context, upload / step / download strictly in sequence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    domain_n, boxes = grid_of_boxes(n_gpus, args.size)
    config = {
        "workload": f"3D spherical Riemann problem, PerfectGas gamma=1.4, "
                    f"MUSCL-Hancock<HLLEM-Larrouturou,VanLeer>, Strang split sweeps, "
                    f"{domain_n[0]}x{domain_n[1]}x{domain_n[2]} cells = {n_gpus} box(es) of "
                    f"{args.size}^3, gcw {GCW[0]}/{GCW[1]}, transmissive, CFL {CFL}",
        "cells_per_gpu": args.size ** 3, "sweeps_per_step": SWEEPS_PER_STEP,
        "l2_policy": "inputs larger than L2 (8.2 GB state per sweep at 512^3)",
        "partition": f"{n_gpus} box(es), one per GPU",
    }

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, el, done = cpu_reference_run(args.steps, min(args.warmup, 1), budget_s=120.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": n_gpus, "steps": done, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * el / max(done, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": base,
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
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == n_gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    dx = [2.0 / domain_n[0]] * 3
    origin = [-1.0, -1.0 * domain_n[1] / domain_n[0], -1.0 * domain_n[2] / domain_n[0]]
    owner = list(range(n_gpus)) if world > 1 else [0] * n_gpus
    ctx = capi.Context(3, [0] * 3, [d - 1 for d in domain_n], dx, boxes, scratch_gcw=GCW[0],
                       flux_gcw=GCW[1], device=local_rank, comm_rank=rank if world > 1 else 0,
                       comm_size=world, owner=owner)
    if world > 1:
        idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()))
    for d in range(3):
        for s in (0, 1):
            ctx.add_bc(d, s, capi.BC_TRANSMISSIVE)

    arith = {"ieee": capi.ARITH_IEEE, "tolerance": capi.ARITH_TOLERANCE}
    other_name = "tolerance" if args.arithmetic == "ieee" else "ieee"
    ctx.set_arithmetic(arith[args.arithmetic])
    config["arithmetic"] = (
        "ieee: every fp64 operation correctly rounded, no FMA contraction, bit-identical to the "
        "CPU oracle" if args.arithmetic == "ieee" else
        "tolerance: reciprocal-multiply quotients + FMA contraction, within 1e-12 relative per "
        "step of the CPU oracle (tests/test_gpu_tolerance_arithmetic.py)")

    my_boxes = ctx.owned_boxes()
    # pinned host copies of the initial state (also the e2e staging buffers)
    host = {}
    for b in my_boxes:
        lo, hi, nc = ctx.state_box(b)
        shape = ctx.state_shape(b)
        t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        box_state(lo, hi, dx, origin, out=t.numpy())
        host[b] = t
        ctx.upload_state(b, t.numpy())
    ctx.fill_ghost_layer()

    stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        dt = CFL * ctx.compute_stable_dt_all()
        ok, fail, _ = ctx.advance_hierarchy(dt, capi.SPLIT_STRANG)
        if not ok:  # CFL retry, RunSimulation.hpp:141-151
            ok, fail, _ = ctx.advance_hierarchy(CFL * fail, capi.SPLIT_STRANG)
            assert ok

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.enable_timing(True)
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        one_step()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - launches0
    sweep_ms, n_sweeps = ctx.sweep_time_ms()
    ctx.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    total_cells = n_gpus * args.size ** 3 if world > 1 else len(my_boxes) * args.size ** 3
    updates = args.steps * SWEEPS_PER_STEP * total_cells
    value = updates / (elapsed_ms * 1e-3)

    # ---- the same steps with the other arithmetic (secondary, not the headline) ----
    ctx.set_arithmetic(arith[other_name])
    for _ in range(3):
        one_step()
    ctx.enable_timing(True)
    s0, n0 = ctx.sweep_time_ms()
    ov0, ov1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ov0.record(stream)
    for _ in range(args.steps):
        one_step()
    ov1.record(stream)
    barrier()
    other_ms = ov0.elapsed_time(ov1)
    s1, n1 = ctx.sweep_time_ms()
    ctx.enable_timing(False)
    if dist is not None:
        t = torch.tensor([other_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other_ms = float(t.item())
    other_sweep_s = (s1 - s0) / max(n1 - n0, 1) * 1e-3
    ctx.set_arithmetic(arith[args.arithmetic])

    # ---- end to end through the C ABI with host buffers ----------------------
    # Every e2e step is one independent batch: the pinned level data of the
    # rank's boxes (valid cells, IntegratorContext::CopyDataToScratch) go to
    # the device, ghost fill + one coarse step (dt pass and six sweeps) run
    # there, the advanced level data (CopyScratchToData) come back to a pinned
    # output buffer.  Batches are pipelined over three contexts (upload of batch k+1,
    # compute of batch k and download of batch k-1 overlap: PCIe is full
    # duplex), each context driven through the stream-ordered C-ABI calls.
    e2e_steps = max(args.e2e_steps, 1)
    g = GCW[0]
    data_in = {b: host[b][:, g:-g, g:-g, g:-g].contiguous().pin_memory() for b in my_boxes}
    bytes_state = sum(data_in[b].numel() * 8 for b in my_boxes)
    # the pipeline needs 3 x 7.5 GB of pinned output buffers per rank on top of
    # the 7.5 GB input; with more than two ranks on one host that exceeds what
    # the host can pin, so those runs stay strictly sequential (one context,
    # the input buffer doubles as output)
    n_pipe = 3 if (args.e2e_pipeline and world <= 2) else 1
    pipe = [ctx]
    for _ in range(n_pipe - 1):
        c = capi.Context(3, [0] * 3, [d - 1 for d in domain_n], dx, boxes, scratch_gcw=GCW[0],
                         flux_gcw=GCW[1], device=local_rank, comm_rank=rank if world > 1 else 0,
                         comm_size=world, owner=owner)
        c.set_arithmetic(arith[args.arithmetic])
        if world > 1:
            idt = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            c.comm_init(bytes(idt.cpu().numpy().tobytes()))
        for d in range(3):
            for s in (0, 1):
                c.add_bc(d, s, capi.BC_TRANSMISSIVE)
        pipe.append(c)
    if n_pipe > 1:
        outs = [{b: torch.empty_like(data_in[b]).pin_memory() for b in my_boxes} for _ in pipe]
    else:
        outs = [data_in]

    def step_of(c):
        dt = CFL * c.compute_stable_dt_all()
        ok, fail, _ = c.advance_hierarchy(dt, capi.SPLIT_STRANG)
        if not ok:
            ok, fail, _ = c.advance_hierarchy(CFL * fail, capi.SPLIT_STRANG)
            assert ok

    def upload(c):
        for b in my_boxes:
            c.copy_data_to_scratch_async(b, data_in[b].numpy())   # H2D from pinned memory

    trace = os.environ.get("FVS_E2E_TRACE")                  # host time line of every batch (debug)
    skip_compute = os.environ.get("FVS_E2E_SKIP_COMPUTE")    # transfers only (debug, not a bench value)

    def run_e2e(nsteps):
        upload(pipe[0])
        t_prev = time.perf_counter()
        for k in range(nsteps):
            cur = pipe[k % n_pipe]
            if k + 1 < nsteps and n_pipe > 1:
                upload(pipe[(k + 1) % n_pipe])               # overlaps cur's compute
            if trace:
                cur.synchronize()
                t_up = time.perf_counter()
            if not skip_compute:
                cur.fill_ghost_layer()
                step_of(cur)                                 # synchronises cur's stream (dt)
            if trace and rank == 0:
                t_now = time.perf_counter()
                print(f"e2e batch {k}: waited for upload {1e3 * (t_up - t_prev):.1f} ms, "
                      f"compute {1e3 * (t_now - t_up):.1f} ms", file=sys.stderr)
                t_prev = t_now
            for b in my_boxes:
                cur.copy_scratch_to_data_async(b, outs[k % n_pipe][b].numpy())   # D2H of the result
            if n_pipe == 1 and k + 1 < nsteps:
                cur.synchronize()
                upload(cur)
        for c in pipe:
            c.synchronize()

    run_e2e(n_pipe)                                          # every context once: buffers allocated
    barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = e2e_steps * SWEEPS_PER_STEP * total_cells / (e2e_ms * 1e-3)
    for c in pipe[1:]:
        c.close()

    # ---- roofline of the dominant kernel (the fused sweep) --------------------
    peak, peak_src = measured_peak_gbs()
    per_launch_bytes = BYTES_PER_CELL_UPDATE * len(my_boxes) * args.size ** 3
    avg_sweep_s = (sweep_ms / max(n_sweeps, 1)) * 1e-3
    achieved = per_launch_bytes / avg_sweep_s / 1e9 if n_sweeps else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("size") == args.size and len(my_boxes) == 1:
                traffic = tj.get("fused_sweep_bytes_per_launch" if args.arithmetic == "ieee"
                                 else "fused_sweep_bytes_per_launch_tolerance")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "kernel": ("fvs::" if args.arithmetic == "ieee" else "fvs_tol::") +
                  "fused_sweep_x_smem_kernel / fused_sweep_yz_smem_kernel<3,*,HLLEM-L,MUSCL,VanLeer>",
        "avg_launch_ms": 1e3 * avg_sweep_s, "launches_timed": n_sweeps,
        "share_of_step": sweep_ms / elapsed_ms if elapsed_ms else None,
        "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
    }

    peak_o = peak
    other = {
        "arithmetic": other_name, "value": updates / (other_ms * 1e-3), "unit": UNIT,
        "ms_per_step": other_ms / args.steps, "avg_launch_ms": 1e3 * other_sweep_s,
        "roofline_frac": (per_launch_bytes / other_sweep_s / 1e9 / peak_o) if other_sweep_s else None,
        "note": "same workload and steps, device-resident, timed right after the headline region",
    }

    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    cpu_baseline = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_run(20, 1, budget_s=15.0)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "clocks": clocks,
        # whole job: every rank moves its own boxes (one 512^3 box per rank)
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_state * max(world, 1),
                "d2h_bytes_per_step": bytes_state * max(world, 1), "steps": e2e_steps,
                "pipeline": f"{n_pipe} context(s); every step is an independent batch "
                            "(pinned level data = valid cells in, ghost fill + dt pass + 6 sweeps on the "
                            "device, level data out: CopyDataToScratch / CopyScratchToData); "
                            "wall clock around all batches"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "other_arithmetic": other,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
