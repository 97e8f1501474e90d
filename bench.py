#!/usr/bin/env python
"""bench.py — Msamples/s of the sample job on BASELINE config 3 (book-1 final scene, BVH +
defocus, 1920x1080, 256 spp, depth 50), one process per GPU.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference          # the reference's CPU algorithm on the host cores

A "step" is one full sample batch of the frame (W*H*spp camera paths).  N > 1: the frame is
sharded by balanced row tiles (total work fixed -> "strong" scaling); every rank's kernel writes its
tile straight into rank 0's frame over NVLink (CUDA IPC mapping; --nccl-gather: the round-1 gather),
and a step ends with a 4-byte all-reduce as the frame-complete signal; time = max over ranks, CUDA
events on the launching stream.  `e2e`: the host-buffer C ABI call with a live cancellation token —
rtb_sample_batch at N = 1, rtb_multi_sample_batch driving all N GPUs from rank 0's process at N > 1.
`value_fast`: the same steps with the opt-in fast-arithmetic build.  The default run (config 3) also
carries short lines for configs 2, 4 and 5 under `other_configs`.  Prints ONE JSON line (rank 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work per unit, SURVEY.md §8(d) (FMA = 2 flop); DESIGN.md "Roofline".
FLOP_SPHERE_TEST = 22
FLOP_NODE_TEST = 24
FLOP_SHADE_STANDARD = 200
FLOP_SHADE_DIELECTRIC = 80
FLOP_SKY = 6
FLOP_CAMERA_RAY = 60
BYTES_PER_PIXEL = 44 + 48      # accumulators in + out (SampleBatchJob.cs:41-49), diagnostics extra

CONFIGS = {
    # name: scene, max_bvh_depth, W, H, spp, trace_depth, aperture
    "c1": ("three_spheres", 0, 400, 225, 4, 8, None),
    "c2": ("final", 0, 1280, 720, 64, 50, None),
    "c3": ("final", 16, 1920, 1080, 256, 50, 0.1),
    "c4": ("final", 16, 3840, 2160, 1024, 50, 0.1),
    "c5": ("stress", 16, 4096, 4096, 2048, 50, 0.1),
    "c5s": ("stress", 16, 1920, 1080, 64, 50, 0.1),     # config 5's world at a single-GPU size
    "mesh": ("mesh", 16, 1920, 1080, 64, 50, 0.1),       # triangle-mesh world (what the reference's host ingests at HEAD)
    "cornell": ("cornell", 16, 1920, 1080, 64, 50, 0.0),  # Rect / Box / moving entities behind rotated transforms
    "fog": ("fog", 16, 1920, 1080, 64, 50, 0.0),          # the same room with participating media (ProbabilisticVolume)
}
STRESS_SPHERES = 10000


def make_scene(host, cfg):
    name, depth = CONFIGS[cfg][0], CONFIGS[cfg][1]
    if name == "mesh":
        return host.make_mesh_scene(max_bvh_depth=depth, subdivisions=4)
    if name == "cornell":
        return host.make_cornell_scene(max_bvh_depth=depth)
    if name == "fog":
        return host.make_cornell_scene(max_bvh_depth=depth, fog=True)
    return host.make_scene(name, max_bvh_depth=depth, target_count=STRESS_SPHERES if name == "stress" else 0)

WORKLOADS = {
    "c1": "three-sphere scene 400x225x4spp depth 8 (linear list)",
    "c2": "book-1 final scene (482 spheres, linear hit list) 1280x720x64spp depth 50",
    "c3": "book-1 final scene (482 spheres) BVH(maxDepth 16) + defocus(aperture 0.1) 1920x1080x256spp depth 50",
    "c4": "book-1 final scene BVH + defocus 3840x2160x1024spp depth 50",
    "c5": "10k-sphere synthetic stress scene BVH(maxDepth 16) + defocus 4096x4096x2048spp depth 50",
    "c5s": "10k-sphere synthetic stress scene BVH(maxDepth 16) + defocus 1920x1080x64spp depth 50",
    "fog": "Cornell box with participating media (4 ProbabilisticVolume balls, one moving; collect-all volume kernel) BVH(maxDepth 16) 1920x1080x64spp depth 50",
    "cornell": "Cornell box of placed entities (6 Rects, 3 Boxes, 2 spheres; one sphere and one box moving) BVH(maxDepth 16) 1920x1080x64spp depth 50",
    "mesh": "triangle-mesh world (5120-triangle smooth icosphere + flat-shaded solids + 3 spheres, 5135 triangles) BVH(maxDepth 16) + defocus 1920x1080x64spp depth 50",
}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.1)] or [r for _, r in self.rows]
        sm, reasons, smax, power = [], set(), None, []
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_reference(cfg, threads, target_seconds=12.0, steps=1, warmup=0):
    """Times the reference's CPU algorithm (our restatement, oracle/liboracle_fast.so: -O3 +
    unsafe-math ≙ Burst FloatMode.Fast, the reference's own xorshift32 stream, one pixel per work
    item ≙ Schedule(W*H, 1)) on a bounded, frame-representative sample: every k-th row of the
    frame at full spp.  Returns (msamples_per_s list per step, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    name, depth, W, H, spp, td, ap = CONFIGS[cfg]
    scene = make_scene(O.rtb.host, cfg)
    # calibrate on 2 rows, then pick the row stride for ~target_seconds
    p = O.rtb.host.make_params(scene, W, H, spp, td, aperture=ap, slice_offset=H // 2 % max(H // 2, 1), slice_divider=max(H // 2, 1))
    buf = O.Buffers(W, H, diagnostics=True)
    t = time.perf_counter()
    O.sample_batch(scene, p, buf, noise=O.NOISE_XORSHIFT, threads=threads, fast=True)
    dt = time.perf_counter() - t
    rows_cal = len(range(p.slice_offset, H, p.slice_divider))
    per_row = dt / rows_cal
    n_rows = int(min(H, max(2, target_seconds / max(per_row, 1e-9))))
    divider = max(1, H // n_rows)
    offset = divider // 2
    p = O.rtb.host.make_params(scene, W, H, spp, td, aperture=ap, slice_offset=offset, slice_divider=divider)
    rows = len(range(offset, H, divider))
    vals = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        O.sample_batch(scene, p, buf, noise=O.NOISE_XORSHIFT, threads=threads, fast=True)
        dt = time.perf_counter() - t
        if i >= warmup:
            vals.append(W * rows * spp / dt / 1e6)
    desc = (f"rows r % {divider} == {offset} of the {W}x{H} frame ({rows} rows = {W * rows * spp / 1e6:.1f} Msamples at {spp} spp), "
            f"{threads} threads, 1 pixel per work item")
    # what the reference's collect-all traversal executes per ray on this sample (its FULL_DIAGNOSTICS counters,
    # Raytracer.cs:56-60; SampleBatchJob.cs:203,428,439), to set beside the pruned walk's executed counts
    d = buf.diagnostics
    rays = float(d["ray_count"].astype("float64").sum())
    cpu_reference.work = {"rays": rays, "boxes_hit_per_ray": float(d["bounds_hit_count"].astype("float64").sum()) / max(rays, 1.0),
                          "candidates_per_ray": float(d["candidate_count"].astype("float64").sum()) / max(rays, 1.0)}
    return vals, desc


def parity_sample(cfg, frame_color4, frame_diag, threads, pixels=6144):
    """The checker beside the measurement: the strict oracle build with the Philox slots (the arithmetic contract both sides
    share) renders `pixels` pixels of the SAME full-size frame — three runs of consecutive pixels in the lower, middle and
    upper third — and is compared with what the timed kernel produced there.  Returns the `parity` object of the JSON line."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as O

    name, depth, W, H, spp, td, ap = CONFIGS[cfg]
    scene = make_scene(O.rtb.host, cfg)
    p = O.rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    run = pixels // 3
    worst, counts_equal, rays_equal, over = 0.0, True, True, 0
    ref = O.Buffers(W, H)
    for row in (H // 6, H // 2, (5 * H) // 6):
        lo = row * W + W // 3
        O.sample_batch(scene, p, ref, noise=O.NOISE_PHILOX, threads=threads, index_range=(lo, lo + run))
        r, g = ref.out_color[lo:lo + run], frame_color4[lo:lo + run]
        counts_equal = counts_equal and bool(np.array_equal(r[:, 3], g[:, 3]))
        rays_equal = rays_equal and bool(np.array_equal(ref.diagnostics["ray_count"][lo:lo + run], frame_diag[lo:lo + run]))
        n = np.maximum(r[:, 3:4], 1)
        d = np.abs(r[:, :3] / n - g[:, :3] / n)
        worst = max(worst, float(d.max()))
        over += int((d.max(axis=1) > 1e-4).sum())
    return {"against": "CPU oracle, strict build, Philox slots (same arithmetic contract)", "pixels": 3 * run, "samples_per_pixel": spp,
            "max_abs_rgb_diff": worst, "tolerance": 1e-4, "pixels_over_tolerance": over,
            "sample_counts_equal": counts_equal, "ray_counts_equal": rays_equal}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    budget = 150.0 / (steps + args.warmup)         # whole run within a few minutes
    vals, desc = cpu_reference(args.config, threads, target_seconds=min(20.0, budget), steps=steps, warmup=args.warmup)
    name, depth, W, H, spp, td, ap = CONFIGS[args.config]
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "Msamples/sec", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (W * H * spp / 1e6) / v, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "note": "ms_per_step extrapolated from the sample to the full frame"},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": desc,
                         "reference_traversal_per_ray": getattr(cpu_reference, "work", None)},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def _pinned_host_buffers(rtb, abi, W, H):
    """plugin.HostBuffers over pinned (page-locked) torch CPU tensors: the host's pooled NativeArrays (Raytracer.cs:279-303)."""
    import torch
    n = W * H
    host = {}
    for k, c in (("color", 4), ("weight", 1), ("normal", 3), ("albedo", 3)):
        host["in_" + k] = torch.zeros(n, c, dtype=torch.float32).pin_memory()
        host["out_" + k] = torch.zeros(n, c, dtype=torch.float32).pin_memory()
    host["diag"] = torch.zeros(n, 4, dtype=torch.float32).pin_memory()
    hb = rtb.plugin.HostBuffers(1, 1)
    hb.width, hb.height = W, H
    hb.in_color, hb.in_weight = host["in_color"].numpy(), host["in_weight"].numpy().reshape(-1)
    hb.in_normal, hb.in_albedo = host["in_normal"].numpy(), host["in_albedo"].numpy()
    hb.out_color, hb.out_weight = host["out_color"].numpy(), host["out_weight"].numpy().reshape(-1)
    hb.out_normal, hb.out_albedo = host["out_normal"].numpy(), host["out_albedo"].numpy()
    hb.diagnostics = host["diag"].numpy().view(abi.DIAGNOSTICS_DTYPE).reshape(-1)
    return hb, host


def measure_config(cfg, steps, warmup, rtb, renderer_mod, sharding, abi, dev, world, rank, local_rank, args, full):
    """Times one BASELINE config at this world size.  `full`: the headline config — also the fast-arithmetic build, the
    instrumented counters for the roofline, the end-to-end leg and the clocks; otherwise a short line (device-resident + e2e)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    name, depth, W, H, spp, td, ap = CONFIGS[cfg]
    scene = make_scene(rtb.host, cfg)
    params = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    fr = renderer_mod.FrameRenderer(scene, W, H, device_index=local_rank)
    samples_per_step = W * H * spp
    n = W * H
    out = {"workload": WORKLOADS[cfg]}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        """Host-side wait (gloo): an NCCL barrier would leave a spinning kernel on every other GPU, and a GPU time-slices
        between processes — rank 0's end-to-end leg drives ALL the GPUs from its own process."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=args.cpu_group)

    assembly = "single GPU"
    if world > 1:
        peer = (not args.nccl_gather) and fr.enable_peer_frame()
        assembly = ("peer frame: every rank's kernel reads and writes its row tile in rank 0's HBM over NVLink (CUDA IPC mapping); "
                    "one 4-byte all-reduce per frame as the frame-complete signal") if peer else \
                   "one batched NCCL gather of all tile buffers to rank 0 per frame"

    # ---- load-balanced row tiles (one process per GPU: the Python twin of the plugin's own balancer, rtb_balance_rows) ----
    # (1) cost model from a cheap instrumented probe batch: per-row executed box tests, entity tests and rays (the
    #     reference's FULL_DIAGNOSTICS counters, Raytracer.cs:56-60) weighted by their instruction cost;
    # (2) feedback from measured per-rank kernel times during the warm-up steps.
    tiles_kind = "equal"
    row_cost = None
    if world > 1 and not args.equal_tiles:
        probe = rtb.host.make_params(scene, W, H, max(1, min(8, spp)), td, aperture=ap, seed=12345)
        fr.ctx.set_option(abi.OPT_COUNTERS, 1)
        fr.render_device(probe, gather=False)
        fr.ctx.set_option(abi.OPT_COUNTERS, 0)
        if fr.peer is None:
            fr.gather(all_ranks=True)
        else:
            fr.frame_complete()
        torch.cuda.synchronize()
        cost_t = torch.zeros(H, dtype=torch.float64, device=dev)
        if fr.peer is None or rank == 0:
            d = fr.diag.view(H, W, 4).double()
            cost_t = (450.0 * d[:, :, 0] + 35.0 * d[:, :, 1] + 40.0 * d[:, :, 2]).sum(dim=1)
        if fr.peer is not None:
            dist.broadcast(cost_t, src=0)
        row_cost = cost_t.cpu().numpy()
        bb = rtb.plugin.balance_rows(row_cost, 0, H, world)
        fr.set_tiles(list(zip(bb[:-1], bb[1:])))
        tiles_kind = "cost-model balanced + kernel-time feedback"

    last_kernel_ms = [0.0]

    def rebalance_from_times():
        """Scale each tile's rows by measured time / modelled cost and re-partition (all ranks compute the same tiles)."""
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = last_kernel_ms[0]
        dist.all_reduce(t)
        times = t.cpu().numpy()
        for g, (b, e) in enumerate(fr.tiles):
            c = row_cost[b:e].sum()
            if c > 0 and times[g] > 0:
                row_cost[b:e] *= times[g] / c
        bb = rtb.plugin.balance_rows(row_cost, 0, H, world)
        fr.set_tiles(list(zip(bb[:-1], bb[1:])))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed_device_steps(k, w, rebalance):
        """w warm-up + k timed device-resident steps; returns (total ms max over ranks, kernel ms of this rank, per rank)."""
        wev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for _ in range(w):
            wev[0].record()
            fr.render_device(params, gather=False)
            wev[1].record()
            fr.frame_complete()
            torch.cuda.synchronize()
            last_kernel_ms[0] = wev[0].elapsed_time(wev[1])
            if rebalance and row_cost is not None:
                rebalance_from_times()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        ev[0].record()
        for i in range(k):
            flush.zero_()                                   # evict L2 between timed iterations
            kev[i][0].record()
            fr.render_device(params, gather=False)
            kev[i][1].record()
            fr.frame_complete()                             # peer frame: 4-byte all-reduce; fallback: the NCCL gather
        ev[1].record()
        barrier()
        total_ms = sharding.max_over_ranks(ev[0].elapsed_time(ev[1]), dev)
        kernel_ms = sum(x.elapsed_time(y) for x, y in kev) / k
        per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
        per_rank[rank] = kernel_ms
        if world > 1:
            dist.all_reduce(per_rank)
        return total_ms, kernel_ms, [round(float(x), 3) for x in per_rank.cpu()]

    # ---- work counters of one step (instrumented kernel, untimed) for the roofline numerator ---
    cnt = None
    if full:
        fr.ctx.set_option(abi.OPT_COUNTERS, 1)
        fr.render_device(params, gather=False)
        cnt = fr.ctx.counters()
        fr.ctx.set_option(abi.OPT_COUNTERS, 0)
        fr.frame_complete()
        torch.cuda.synchronize()

    # ---- device-resident throughput (the parity build: the headline) --------------------------------
    sampler = ClockSampler(local_rank) if (rank == 0 and full) else None
    t0 = time.perf_counter()
    total_ms, kernel_ms, kernel_ms_per_rank = timed_device_steps(steps, warmup, True)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if sampler else None
    out.update(value=samples_per_step * steps / (total_ms * 1e-3) / 1e6, ms_per_step=total_ms / steps, kernel_ms=kernel_ms,
               kernel_ms_max=max(kernel_ms_per_rank), kernel_ms_per_rank=kernel_ms_per_rank, clocks=clocks, counters=cnt,
               row_tiles=[list(map(int, t)) for t in fr.tiles], tiles_kind=tiles_kind, assembly=assembly)
    if rank == 0:
        out["frame_checksum"] = float(fr.out["color"][:, :3].double().sum().item())
    barrier()       # peer frame: the other ranks must not start the next image while rank 0 reads this one

    # ---- the opt-in fast-arithmetic build, same steps (RTB_OPT_MATH = 1; statistics: tools/fast_math_report.py) ----
    if full:
        fr.ctx.set_option(abi.OPT_MATH, abi.MATH_FAST)
        f_total, f_kernel, _ = timed_device_steps(steps, 1, False)
        fr.ctx.set_option(abi.OPT_MATH, abi.MATH_PARITY)
        out["value_fast"] = samples_per_step * steps / (f_total * 1e-3) / 1e6
        out["ms_per_step_fast"] = f_total / steps
        fast = fr.out["color"].clone() if rank == 0 else None
        barrier()
        if rank == 0:
            fr.render_device(params, gather=False)
            fr.frame_complete()
            torch.cuda.synchronize()
            strict = fr.out["color"]
            cs, cf = strict[:, 3:4].clamp(min=1), fast[:, 3:4].clamp(min=1)
            d = (strict[:, :3] / cs - fast[:, :3] / cf).abs().max(dim=1).values
            out["fast_vs_parity"] = {"mean_abs_rgb_diff": float(d.mean()), "rmse": float((d * d).mean().sqrt()),
                                     "p999_abs_rgb_diff": float(torch.quantile(d[:: max(1, d.numel() // 1000000)].float(), 0.999)),
                                     "pixels_over_1e-4": float((d > 1e-4).float().mean()),
                                     "pixels_with_other_sample_count": float((strict[:, 3] != fast[:, 3]).float().mean()),
                                     "image_mean_diff": float((strict[:, :3] / cs).mean() - (fast[:, :3] / cf).mean())}
        else:
            fr.render_device(params, gather=False)
            fr.frame_complete()
        barrier()

    # ---- end to end through the host-buffer C ABI, with a LIVE cancellation token (the C# job always passes one) -------
    # N = 1: rtb_sample_batch.  N > 1: ONE process (rank 0) drives all N GPUs through rtb_multi_sample_batch — the reference's
    # single call site (Raytracer.cs:671-736) — on one set of pinned host arrays every device reads and writes in place; the
    # other ranks only wait.  The timed region holds the host<->device traffic of every step (in place over PCIe).
    cancel = np.zeros(1, np.uint8)
    e2e = None
    cpu_barrier()
    if rank == 0:
        hb, host = _pinned_host_buffers(rtb, abi, W, H)
        if world == 1:
            ctx_e2e, api = fr.ctx, "rtb_sample_batch (C ABI, pinned host buffers read and written in place by the kernel over PCIe, live cancellation token)"

            def e2e_step():
                ctx_e2e.sample_batch(params, hb, cancel=cancel)
        else:
            multi = rtb.plugin.MultiContext(list(range(world)))
            multi.upload(scene)
            api = (f"rtb_multi_sample_batch (C ABI, ONE process drives {world} GPUs: row tiles balanced inside the plugin, every device reads and "
                   "writes its tile of the pinned host frame in place over its own PCIe link, live cancellation token; no gather)")

            def e2e_step():
                multi.sample_batch(params, hb, cancel=cancel)
        for _ in range(max(2, min(warmup, 3)) if full else 1):
            e2e_step()
        t = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        e2e_s = time.perf_counter() - t
        # the same loop without a token: the token must cost nothing
        if full:
            t = time.perf_counter()
            for _ in range(steps):
                (ctx_e2e.sample_batch(params, hb) if world == 1 else multi.sample_batch(params, hb))
            e2e_s_no_token = time.perf_counter() - t
        e2e = {"value": samples_per_step * steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": n * 44, "d2h_bytes_per_step": n * (48 + 16),
               "api": api, "in_place": bool(fr.ctx.last_batch_in_place()) if world == 1 else True,
               "out_color_checksum": float(host["out_color"][:, :3].double().sum().item())}
        if full:
            e2e["value_without_token"] = samples_per_step * steps / e2e_s_no_token / 1e6
        if world > 1:
            bounds, ms = multi.tiles()
            e2e["row_bounds"], e2e["kernel_ms_per_device"] = bounds, [round(x, 3) for x in ms]
            multi.close()
        out["host_frame"] = (host["out_color"].numpy(), host["diag"].numpy()[:, 0])
    cpu_barrier()
    out["e2e"] = e2e
    fr.close()
    del flush
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rtb = importlib.import_module("raytracing-in-one-weekend_b200")
    renderer_mod = importlib.import_module("raytracing-in-one-weekend_b200.renderer")
    sharding = importlib.import_module("raytracing-in-one-weekend_b200.sharding")
    abi = rtb.abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    args.cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        args.cpu_group = dist.new_group(backend="gloo")

    common = (rtb, renderer_mod, sharding, abi, dev, world, rank, local_rank, args)
    m = measure_config(args.config, args.steps, args.warmup, *common, full=True)
    cnt = m["counters"]
    name, depth, W, H, spp, td, ap = CONFIGS[args.config]
    n = W * H

    # ---- the other BASELINE configs, short runs (they are parity-test cases; their lines ride along) --------------
    others = {}
    if args.config == "c3" and not args.no_other_configs:
        for cfg in ("c2", "c4", "c5"):
            try:
                o = measure_config(cfg, 1 if cfg == "c5" else 2, 1, *common, full=False)
                others[cfg] = {"workload": o["workload"], "value": o["value"], "unit": "Msamples/s", "ms_per_step": o["ms_per_step"],
                               "kernel_ms_per_rank": o["kernel_ms_per_rank"], "e2e": ({k: v for k, v in o["e2e"].items() if k != "api"} if o["e2e"] else None),
                               "frame_checksum": o.get("frame_checksum"), "steps": 1 if cfg == "c5" else 2, "warmup": 1}
            except Exception as e:      # noqa: BLE001 — an extra line must not take the headline down
                others[cfg] = {"error": str(e)[:300]}

    # ---- CPU baseline + parity object (rank 0, N = 1 only) -------------------------------------------------------
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        vals, desc = cpu_reference(args.config, threads, target_seconds=12.0)
        cpu = {"value": vals[0], "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": desc,
               "reference_traversal_per_ray": getattr(cpu_reference, "work", None)}
        try:
            parity = parity_sample(args.config, *m["host_frame"], threads)
        except Exception as e:      # noqa: BLE001 — the checker must not take the measurement down
            parity = {"error": str(e)}

    if rank == 0:
        fp32_peak = None
        try:
            ctx = rtb.plugin.Context(local_rank)
            fp32_peak = ctx.measure_fp32_peak(5)
            ctx.close()
        except Exception:       # noqa: BLE001
            pass
        traffic, traffic_source = None, None
        summary = os.path.join(ROOT, "profiles", "ncu_summary_latest.json")
        if os.path.exists(summary):
            try:
                js = json.load(open(summary))
                traffic = js.get("dram_bytes_per_launch")
                traffic_source = f"profiles/ncu_summary_latest.json ({js.get('label', '')}; ncu --set full capture of {js.get('captured', 'this round')}, not of this run)"
            except (OSError, ValueError):
                traffic = None
        kernel_ms = m["kernel_ms"]
        flops_rank = (cnt["sphere_tests"] * FLOP_SPHERE_TEST + cnt["node_tests"] * FLOP_NODE_TEST
                      + cnt["shade_standard"] * FLOP_SHADE_STANDARD + cnt["shade_dielectric"] * FLOP_SHADE_DIELECTRIC
                      + cnt["sky_hits"] * FLOP_SKY + cnt["samples"] * FLOP_CAMERA_RAY)
        achieved_tf = flops_rank / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": "Msamples/sec", "value": m["value"], "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOADS[args.config], "seed": 1, "rng": "Philox4x32-10 keyed (pixel, sample, bounce)",
                "parallelism": f"row tiles x{world} ({m['tiles_kind']}); {m['assembly']}" if world > 1 else "single GPU",
                "l2": "256 MB buffer written between timed iterations (inside the bracket, ~0.05 ms) and 191 MB of accumulators per step > 126 MB L2",
                "outputs": "color + normal + albedo + sampleCountWeight + diagnostics (full job contract)",
                "arithmetic": "parity build (strict IEEE, FMAs only where written; bit-checked against the CPU restatement); value_fast = the opt-in fast build",
                "tree": "RTB_OPT_RETREE = 1 (plugin default): the walk's topology is a surface-area-heuristic tree over the host BVH's leaves — same leaf boxes, same candidates, bit-identical frame (checksum equal to the host topology's)",
            },
            "e2e": m["e2e"],
            "value_fast": m.get("value_fast"), "ms_per_step_fast": m.get("ms_per_step_fast"), "fast_vs_parity": m.get("fast_vs_parity"),
            "gpu_launches": args.steps * world,
            "kernel_ms_rank0": kernel_ms, "kernel_ms_max_rank": m["kernel_ms_max"], "kernel_ms_per_rank": m["kernel_ms_per_rank"],
            "row_tiles": m["row_tiles"], "frame_checksum": m.get("frame_checksum"),
            "mrays_per_s": cnt["rays"] / (kernel_ms * 1e-3) / 1e6 if world == 1 else None,
            "failed_sample_fraction": cnt["failed_samples"] / max(cnt["samples"], 1),
            "roofline": {
                "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak if fp32_peak else None,
                "traffic": traffic, "traffic_source": traffic_source,
                "note": "path is FP32-ALU/latency bound (SURVEY §8d): executed algorithmic flops of rank 0's tile (kernel counters x per-unit constants) / rank 0 kernel time; peak = FP32 FMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32-pipe figure)",
                "counters_rank0": cnt,
                "executed_per_ray": {"box_tests": cnt["node_tests"] / max(cnt["rays"], 1), "sphere_tests": cnt["sphere_tests"] / max(cnt["rays"], 1)},
            },
            "roofline_hbm": {
                "bound": "hbm", "achieved": BYTES_PER_PIXEL * n / world / (kernel_ms * 1e-3) / 1e9, "peak": _hbm_peak(), "unit": "GB/s",
                "frac": BYTES_PER_PIXEL * n / world / (kernel_ms * 1e-3) / 1e9 / _hbm_peak(), "traffic": traffic,
                "note": "92 B/pixel/batch algorithmic; HBM is not the bound of this path",
            },
            "clocks": m["clocks"],
        }
        if others:
            line["other_configs"] = others
        if cpu:
            line["cpu_baseline"] = cpu
            line["parity"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        return 6650.0     # B200_PROFILING.md fallback


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # NCCL / torchrun banners must not pollute the one-line contract
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--equal-tiles", action="store_true", help="contiguous equal row tiles instead of ray-count balanced ones")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: assemble the frame with the NCCL gather instead of the peer-mapped frame on rank 0")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short c2 / c4 / c5 lines that ride along with the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
