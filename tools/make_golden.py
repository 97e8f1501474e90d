"""Generates tests/golden/*.npz with the CPU oracle (strict build, Philox slots).

The reference has no golden vectors for this path and cannot run here, so
these fixtures pin OUR oracle against regressions and give the GPU tests a target that does
not need the oracle at run time.  Regenerate only when the oracle's contract changes:
    python tools/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402

CASES = {
    # name: (scene, max_bvh_depth, W, H, spp, trace_depth, aperture, noise)
    "three_spheres_32x18x4_d8_philox": ("three_spheres", 0, 32, 18, 4, 8, None, O.NOISE_PHILOX),
    "final_linear_32x18x4_d50_philox": ("final", 0, 32, 18, 4, 50, None, O.NOISE_PHILOX),
    "final_bvh16_defocus_48x27x8_d50_philox": ("final", 16, 48, 27, 8, 50, 0.1, O.NOISE_PHILOX),
    "three_spheres_32x18x4_d8_xorshift": ("three_spheres", 0, 32, 18, 4, 8, None, O.NOISE_XORSHIFT),
    "mesh_bvh16_48x27x8_d50_philox": ("mesh", 16, 48, 27, 8, 50, None, O.NOISE_PHILOX),
    "cornell_bvh16_48x27x8_d50_philox": ("cornell", 16, 48, 27, 8, 50, None, O.NOISE_PHILOX),        # Rect / Box / moving entities
    "cornell_bvh16_32x18x4_d50_xorshift": ("cornell", 16, 32, 18, 4, 50, None, O.NOISE_XORSHIFT),
    "fog_bvh16_48x27x8_d50_philox": ("fog", 16, 48, 27, 8, 50, None, O.NOISE_PHILOX),               # ProbabilisticVolume media
    "fog_bvh16_32x18x4_d50_xorshift": ("fog", 16, 32, 18, 4, 50, None, O.NOISE_XORSHIFT),
    "textured_mesh_bvh16_48x27x8_d50_philox": ("textured_mesh", 16, 48, 27, 8, 50, None, O.NOISE_PHILOX),   # image textures
}


def make_scene(name, depth):
    if name == "mesh":
        return O.rtb.host.make_mesh_scene(max_bvh_depth=depth)
    if name == "textured_mesh":
        return O.rtb.host.make_mesh_scene(max_bvh_depth=depth, textured=True)
    if name in ("cornell", "fog"):
        return O.rtb.host.make_cornell_scene(max_bvh_depth=depth, fog=(name == "fog"))
    return O.rtb.host.make_scene(name, max_bvh_depth=depth)



def render(case):
    name, depth, W, H, spp, td, ap, noise = CASES[case]
    scene = make_scene(name, depth)
    p = O.rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    b = O.Buffers(W, H)
    O.sample_batch(scene, p, b, noise=noise, threads=1)
    return b


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and case not in only:
            continue
        b = render(case)
        np.savez_compressed(
            os.path.join(out_dir, case + ".npz"), color=b.out_color, normal=b.out_normal, albedo=b.out_albedo,
            weight=b.out_weight, ray_count=b.diagnostics["ray_count"],
        )
        print(case, "mean rgb", b.rgb().mean(axis=(0, 1)))


if __name__ == "__main__":
    main()
