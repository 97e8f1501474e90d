"""Small batches through every kernel for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-in-one-weekend_b200")
abi = rtb.abi
ctx = rtb.plugin.Context(0)
cases = [(rtb.host.make_scene("final", max_bvh_depth=16), 0.1, 1), (rtb.host.make_scene("final", max_bvh_depth=16), 0.1, 8),
         (rtb.host.make_mesh_scene(max_bvh_depth=16), 0.0, 1), (rtb.host.make_scene("stress", max_bvh_depth=16, target_count=10000), 0.1, 1)]
wide = os.environ.get("SANITIZE_WIDE", "0") == "1"     # the worlds beyond the BASELINE configs instead
if wide:
    cases = [(rtb.host.make_cornell_scene(max_bvh_depth=16), 0.0, 1), (rtb.host.make_cornell_scene(max_bvh_depth=2, fog=True), 0.1, 1),
             (rtb.host.make_cornell_scene(max_bvh_depth=16, fog=True), 0.0, 1),
             (rtb.host.make_mesh_scene(max_bvh_depth=16, textured=True), 0.0, 1), (rtb.host.make_random_placed_scene(3, count=24), 0.1, 1)]
W, H, spp = 64, 36, 8
for scene, ap, k in cases:
    ctx.set_option(abi.OPT_LEAF_SPHERES, k)
    ctx.upload(scene)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=ap)
    for kernel in (abi.KERNEL_SIMPLE, abi.KERNEL_MEGA):
        for counters in (0, 1):
            ctx.set_option(abi.OPT_KERNEL, kernel)
            ctx.set_option(abi.OPT_COUNTERS, counters)
            b = rtb.plugin.HostBuffers(W, H)
            ctx.sample_batch(p, b)
            if wide and kernel == abi.KERNEL_SIMPLE and not counters:      # the white-noise stream through the same kernels
                ctx.set_option(abi.OPT_NOISE, abi.NOISE_WHITE)
                ctx.sample_batch(p, rtb.plugin.HostBuffers(W, H))
                ctx.set_option(abi.OPT_NOISE, abi.NOISE_PHILOX)
            print(scene.name, "leaf", k, "kernel", kernel, "counters", counters, float(b.out_color[:, :3].sum()), flush=True)
ctx.close()
print("done")
