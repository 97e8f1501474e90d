"""Kernel time of config 3 rendered as one launch vs as two row bands (per-launch overhead / tail check)."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-in-one-weekend_b200")
W, H, spp = 1920, 1080, 256
scene = rtb.host.make_scene("final", max_bvh_depth=16)
ctx = rtb.plugin.Context(0); ctx.upload(scene)
b = rtb.plugin.HostBuffers(W, H); ctx.register_host_buffers(b)
def ms(r0, r1):
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, row_begin=r0, row_end=r1)
    best = 1e9
    for _ in range(3):
        ctx.sample_batch(p, b); best = min(best, ctx.last_kernel_ms())
    return best
split = int(sys.argv[1]) if len(sys.argv) > 1 else 414
res = {"full": ms(0, H), "low": ms(0, split), "high": ms(split, H)}
res["sum"] = res["low"] + res["high"]
eighths = [ms(i * H // 8, (i + 1) * H // 8) for i in range(8)]
res["eighths"] = [round(x, 2) for x in eighths]; res["eighths_sum"] = sum(eighths)
print(json.dumps(res))
