"""A few launches of the finalize kernel (FinalizeTexturesJob on device buffers) at 3840x2160 for ncu captures."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

rtb = importlib.import_module("raytracing-in-one-weekend_b200")
W, H = 3840, 2160
n = W * H
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(3)
col = (torch.rand(n, 3, generator=g) * 1.6 - 0.2).to(dev)
nor = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1).to(dev)
alb = torch.rand(n, 3, generator=g).to(dev)
out = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(3)]
ctx = rtb.plugin.Context(0)
for _ in range(3):
    ctx.finalize_device(W, H, col, nor, alb, out[0], out[1], out[2], stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("finalized", int(out[0].sum().item()) & 0xffff)
