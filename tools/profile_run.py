"""One device-resident step of BASELINE config 3 (or --config) for ncu captures."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

sys.argv = [sys.argv[0]] + sys.argv[1:]
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--shrink", type=int, default=1, help="divide width, height and spp by this (short captures of slow kernels)")
args = ap.parse_args()
bench = importlib.import_module("bench")
rtb = importlib.import_module("raytracing-in-one-weekend_b200")
renderer = importlib.import_module("raytracing-in-one-weekend_b200.renderer")
name, depth, W, H, spp, td, aperture = bench.CONFIGS[args.config]
W, H, spp = W // args.shrink, H // args.shrink, max(1, spp // args.shrink)
scene = bench.make_scene(rtb.host, args.config)
p = rtb.host.make_params(scene, W, H, spp, td, aperture=aperture)
fr = renderer.FrameRenderer(scene, W, H, 0)
for _ in range(args.steps):
    fr.render_device(p)
torch.cuda.synchronize()
print("rendered", args.config, float(fr.out["color"][:, :3].sum().item()))
