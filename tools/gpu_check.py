"""Developer check run on the GPU box: GPU kernels vs the CPU oracle on small cases, then a
timing of the megakernel on larger ones.  (Uses the oracle as checker only.)"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402

rtb = O.rtb
abi = rtb.abi


def compare(name, a, b):
    rgb_a, rgb_b = a.rgb(), b.rgb()
    d = np.abs(rgb_a - rgb_b)
    out = {
        "case": name,
        "max_abs_rgb": float(d.max()), "mean_abs_rgb": float(d.mean()), "pixels_gt_1e-4": int((d.max(axis=2) > 1e-4).sum()),
        "count_equal": bool(np.array_equal(a.out_color[:, 3], b.out_color[:, 3])),
        "rays_equal": bool(np.array_equal(a.diagnostics["ray_count"], b.diagnostics["ray_count"])),
        "color_bit_equal": bool(np.array_equal(a.out_color, b.out_color)),
        "max_abs_normal": float(np.abs(a.out_normal - b.out_normal).max()),
        "max_abs_albedo": float(np.abs(a.out_albedo - b.out_albedo).max()),
        "max_abs_weight": float(np.abs(a.out_weight - b.out_weight).max()),
    }
    print(json.dumps(out), flush=True)
    return out


def main():
    results = []
    ctx = rtb.plugin.Context(0)
    cases = [
        ("three_spheres", 0, 400, 225, 4, 8, None),
        ("final", 0, 160, 90, 8, 50, None),
        ("final", 16, 160, 90, 8, 50, 0.1),
        ("final", 16, 320, 180, 32, 50, 0.1),
    ]
    for name, depth, W, H, spp, td, ap in cases:
        scene = rtb.host.make_scene(name, max_bvh_depth=depth)
        p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
        ref = O.Buffers(W, H)
        t = time.time()
        O.sample_batch(scene, p, ref, noise=O.NOISE_PHILOX)
        t_cpu = time.time() - t
        ctx.upload(scene)
        for kernel in (abi.KERNEL_SIMPLE, abi.KERNEL_MEGA):
            ctx.set_option(abi.OPT_KERNEL, kernel)
            got = rtb.plugin.HostBuffers(W, H)
            ctx.sample_batch(p, got)
            r = compare(f"{name}/bvh{depth}/{W}x{H}x{spp}/k{kernel}", ref, got)
            r["kernel_ms"] = ctx.last_kernel_ms()
            r["cpu_s"] = t_cpu
            results.append(r)
    # timing
    ctx.set_option(abi.OPT_KERNEL, abi.KERNEL_MEGA)
    for name, depth, W, H, spp, td, ap in [("final", 0, 1280, 720, 64, 50, None), ("final", 16, 1920, 1080, 256, 50, 0.1)]:
        scene = rtb.host.make_scene(name, max_bvh_depth=depth)
        ctx.upload(scene)
        p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
        got = rtb.plugin.HostBuffers(W, H)
        for it in range(2):
            t = time.time()
            ctx.sample_batch(p, got)
            wall = time.time() - t
            ms = ctx.last_kernel_ms()
            r = {"timing": f"{name}/bvh{depth}/{W}x{H}x{spp}", "kernel_ms": ms, "wall_s": wall,
                 "msamples_per_s": W * H * spp / ms / 1e3, "rays": float(got.diagnostics["ray_count"].astype(np.float64).sum()),
                 "failed": float(W * H * spp - got.out_color[:, 3].astype(np.float64).sum()),
                 "mean_rgb": got.rgb().mean(axis=(0, 1)).tolist()}
            print(json.dumps(r), flush=True)
            results.append(r)
    ctx.set_option(abi.OPT_COUNTERS, 1)
    ctx.sample_batch(p, got)
    c = ctx.counters()
    c["kernel_ms_counters"] = ctx.last_kernel_ms()
    print(json.dumps(c), flush=True)
    results.append(c)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
