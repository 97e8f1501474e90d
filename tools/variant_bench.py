"""Times experiment builds of the plugin (lib/variants/librtb_<tag>.so) on the GPU box, each in its own
process: a parity check against the oracle on a small case, then kernel time on config 3 (and 2).
usage: variant_bench.py [tag ...]   (no tags: every variant found + the default build)"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import oracle_lib as O
rtb = O.rtb; abi = rtb.abi
ctx = rtb.plugin.Context(0)
res = {"tag": %(tag)r}
scene = rtb.host.make_scene("final", max_bvh_depth=16)
p = rtb.host.make_params(scene, 128, 72, 16, 50, aperture=0.1)
ref = O.Buffers(128, 72); O.sample_batch(scene, p, ref)
ctx.upload(scene)
got = rtb.plugin.HostBuffers(128, 72); ctx.sample_batch(p, got)
res["parity_max_rgb"] = float(np.abs(ref.rgb() - got.rgb()).max())
res["counts_equal"] = bool(np.array_equal(ref.out_color[:, 3], got.out_color[:, 3]) and np.array_equal(ref.diagnostics["ray_count"], got.diagnostics["ray_count"]))
res["aov_max"] = float(max(np.abs(ref.out_normal - got.out_normal).max(), np.abs(ref.out_albedo - got.out_albedo).max(), np.abs(ref.out_weight - got.out_weight).max()))
for name, depth, W, H, spp, td, ap, key in [("final", 16, 1920, 1080, 256, 50, 0.1, "c3"), ("final", 0, 1280, 720, 64, 50, None, "c2"),
                                            ("stress", 16, 1920, 1080, 64, 50, 0.1, "c5s"), ("mesh", 16, 1920, 1080, 64, 50, 0.1, "mesh"),
                                            ("cornell", 16, 1920, 1080, 64, 50, 0.0, "cornell"), ("fog", 16, 1920, 1080, 64, 50, 0.0, "fog"),
                                            ("cornell", 0, 1920, 1080, 64, 50, 0.0, "cornell0"), ("cornell", 2, 1920, 1080, 64, 50, 0.0, "cornell2")]:
    if key not in %(cfgs)r: continue
    scene = rtb.host.make_mesh_scene(max_bvh_depth=depth, subdivisions=4) if name == "mesh" else rtb.host.make_cornell_scene(max_bvh_depth=depth, fog=(name == "fog")) if name in ("cornell", "fog") else rtb.host.make_scene(name, max_bvh_depth=depth, target_count=10000 if name == "stress" else 0)
    ctx.upload(scene)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    b = rtb.plugin.HostBuffers(W, H, diagnostics=True)
    ms = []
    for _ in range(3):
        ctx.sample_batch(p, b); ms.append(ctx.last_kernel_ms())
    res[key + "_ms"] = min(ms); res[key + "_msamples"] = W * H * spp / min(ms) / 1e3
    res[key + "_checksum"] = float(b.out_color.astype(np.float64).sum())
    if %(fast)r:
        ctx.set_option(abi.OPT_MATH, abi.MATH_FAST)
        ms = []
        for _ in range(3):
            ctx.sample_batch(p, b); ms.append(ctx.last_kernel_ms())
        ctx.set_option(abi.OPT_MATH, abi.MATH_PARITY)
        res[key + "_fast_ms"] = min(ms); res[key + "_fast_checksum"] = float(b.out_color.astype(np.float64).sum())
print("RESULT " + json.dumps(res))
'''


def main():
    tags = [a for a in sys.argv[1:] if not a.startswith("--")]
    sweep = [a[len("--env="):] for a in sys.argv[1:] if a.startswith("--env=")]   # --env=NAME=v1,v2: default build per value
    cfgs = ["c3"] + (["c2"] if "--c2" in sys.argv else []) + (["c5s"] if "--c5s" in sys.argv else []) + (["mesh"] if "--mesh" in sys.argv else []) + (["cornell"] if "--cornell" in sys.argv else []) + (["fog"] if "--fog" in sys.argv else []) + (["cornell0", "cornell2"] if "--cornell0" in sys.argv else [])
    vdir = os.path.join(ROOT, "raytracing-in-one-weekend_b200", "lib", "variants")
    libs = {"default": None}
    for f in sorted(glob.glob(os.path.join(vdir, "librtb_*.so"))):
        libs[os.path.basename(f)[len("librtb_"):-3]] = f
    if tags:
        libs = {t: libs[t] for t in tags}
    out = []
    runs = []
    for tag, path in libs.items():
        if not sweep:
            runs.append((tag, path, {}))
        for sw in sweep:
            name, vals = sw.split("=", 1)
            runs += [(f"{tag}:{name}={v}", path, {name: v}) for v in vals.split(",")]
    for tag, path, extra in runs:
        env = dict(os.environ)
        env.update(extra)
        if path:
            env["RTB_PLUGIN_LIB"] = path
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "tag": tag, "cfgs": cfgs, "fast": "--fast" in sys.argv}], env=env, capture_output=True, text=True, timeout=900)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        if line:
            out.append(json.loads(line[0][7:]))
            print(line[0][7:], flush=True)
        else:
            print(tag, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
