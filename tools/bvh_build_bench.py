"""How long does the host take to build the reference's BVH (BvhNodeData.cs:122-213, restated in csrc/host/rtb_host.cpp) for
large worlds — the per-world-change cost a GPU builder would remove?  CPU only; prints a markdown table.
usage: bvh_build_bench.py [--json out.json]"""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rtb = importlib.import_module("raytracing-in-one-weekend_b200")
abi = rtb.abi


def build(bounds, max_depth):
    n = len(bounds)
    order = np.zeros(max(n, 1), np.uint32)
    cap = 2 * n + 1
    nodes = np.zeros(cap, dtype=abi.BVH_NODE_DTYPE)
    count = C.c_size_t(0)
    t = time.perf_counter()
    rc = rtb.host.lib().rtbh_build_bvh_from_bounds(bounds.ctypes.data, n, max_depth, order.ctypes.data, n, nodes.ctypes.data, cap, C.byref(count))
    dt = time.perf_counter() - t
    assert rc == 0
    return dt, count.value


def mesh_bounds(n_tri, rng):
    """Bounds of a synthetic triangle soup with mesh-like statistics: small triangles on a few closed surfaces."""
    c = rng.normal(size=(n_tri, 3)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    c *= rng.choice([1.0, 2.5, 4.0], size=(n_tri, 1)).astype(np.float32)
    c += rng.choice([-6.0, 0.0, 6.0], size=(n_tri, 1)).astype(np.float32) * np.array([[1.0, 0.0, 0.3]], np.float32)
    e = (np.abs(rng.normal(size=(n_tri, 3))) * (2.0 / np.sqrt(n_tri)) + 1e-3).astype(np.float32)
    return np.ascontiguousarray(np.hstack([c - e, c + e]), dtype=np.float32)


def main():
    rng = np.random.default_rng(7)
    rows = []
    s = rtb.host.make_scene("stress", max_bvh_depth=16, target_count=10000)
    b = np.zeros((len(s.spheres), 6), np.float32)
    for i in range(len(s.spheres)):
        rtb.host.lib().rtbh_sphere_bounds(s.spheres[i:i + 1].ctypes.data, b[i].ctypes.data)
    for label, bounds, depth in [("10 004 spheres (config 5)", b, 16)] + [
            (f"{n:,} triangles (synthetic mesh)".replace(",", " "), mesh_bounds(n, rng), 32) for n in (81_927, 250_000, 1_000_000)]:
        best = min(build(bounds, depth)[0] for _ in range(3))
        _, nodes = build(bounds, depth)
        rows.append({"world": label, "entities": len(bounds), "max_depth": depth, "nodes": int(nodes), "build_ms": best * 1e3})
    print("| world | entities | maxDepth | nodes | host build |\n|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['world']} | {r['entities']} | {r['max_depth']} | {r['nodes']} | {r['build_ms']:.1f} ms |")
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
