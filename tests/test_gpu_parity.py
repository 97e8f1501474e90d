"""GPU parity tests proper: every call goes through the C ABI (librtb.so) on cuda:0 and is
checked against the CPU oracle on the same seeded inputs, or against the committed golden
fixtures, or — at BASELINE.json's full size — through size-independent properties.

Parity contract:
  * discrete path decisions are bit-identical: per-pixel successful-sample counts (color.w) and
    ray counts are EXACTLY equal;
  * the validation kernel (thread per pixel, reference accumulation order) is bit-identical in
    every output;
  * the megakernel's sums are order-independent fixed point: per-pixel RGB within 1e-4
    (BASELINE.json north_star tolerance; observed < 1e-6).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4       # BASELINE.json north_star: "per-pixel RGB within 1e-4"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def render_gpu(rtb, ctx, scene, p, W, H, kernel, inputs=None):
    ctx.upload(scene)
    ctx.set_option(rtb.abi.OPT_KERNEL, kernel)
    b = rtb.plugin.HostBuffers(W, H)
    if inputs is not None:
        b.in_color[:], b.in_weight[:], b.in_normal[:], b.in_albedo[:] = inputs
    ctx.sample_batch(p, b)
    return b


def assert_parity(ref, got, exact):
    assert np.array_equal(ref.out_color[:, 3], got.out_color[:, 3]), "successful-sample counts differ"
    assert np.array_equal(ref.diagnostics["ray_count"], got.diagnostics["ray_count"]), "ray counts differ"
    if exact:
        assert np.array_equal(ref.out_color, got.out_color)
        assert np.array_equal(ref.out_normal, got.out_normal)
        assert np.array_equal(ref.out_albedo, got.out_albedo)
        assert np.array_equal(ref.out_weight, got.out_weight)
    else:
        assert np.abs(ref.rgb() - got.rgb()).max() <= RGB_TOL
        n = np.maximum(ref.out_color[:, 3:4], 1)
        assert np.abs(ref.out_normal - got.out_normal).max() / n.max() <= RGB_TOL
        assert np.abs(ref.out_albedo / n - got.out_albedo / n).max() <= RGB_TOL
        assert np.abs(ref.out_weight / n[:, 0] - got.out_weight / n[:, 0]).max() <= RGB_TOL
    np.testing.assert_array_equal(np.isnan(ref.diagnostics["sample_count_weight"]), np.isnan(got.diagnostics["sample_count_weight"]))


CASES = [
    # scene, bvh depth, W, H, spp, trace depth, aperture, jitter
    ("three_spheres", 0, 400, 225, 4, 8, None, True),        # BASELINE config 1, full size
    ("three_spheres", 2, 64, 36, 16, 50, 0.2, True),
    ("three_spheres", 0, 33, 17, 5, 8, None, False),          # ragged size, jitter off
    ("final", 0, 96, 54, 4, 50, None, True),                  # linear hit list (config 2 shape)
    ("final", 16, 128, 72, 16, 50, 0.1, True),                # BVH + defocus (config 3 shape)
    ("final", 32, 64, 36, 8, 50, 0.1, True),                  # prefab maxBvhDepth
    ("final", 3, 64, 36, 8, 12, 0.1, True),                   # multi-entity leaves, short paths fail
    ("final", 16, 31, 19, 40, 50, 0.0, True),                 # more samples than a tile, odd size
]


@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-bvh{c[1]}-{c[2]}x{c[3]}x{c[4]}-d{c[5]}")
def test_sample_batch_matches_the_oracle(rtb, oracle, ctx, case, kernel):
    name, depth, W, H, spp, td, ap, jitter = case
    scene = rtb.host.make_scene(name, max_bvh_depth=depth)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap, jitter=jitter)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


@pytest.mark.parametrize("case", ["cornell_bvh16_48x27x8_d50_philox", "fog_bvh16_48x27x8_d50_philox",
                                  "textured_mesh_bvh16_48x27x8_d50_philox", "cornell_bvh16_32x18x4_d50_xorshift",
                                  "fog_bvh16_32x18x4_d50_xorshift"])
def test_gpu_matches_golden_fixtures_of_the_wider_worlds(rtb, ctx, case):
    """The committed fixtures of the worlds beyond the BASELINE configs (placed entities, media, image textures), without the
    oracle at run time: the per-pixel / volume kernel bit for bit in both noise modes, the megakernel's decisions and RGB."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name, depth, W, H, spp, td, ap, noise = mg.CASES[case]
    scene = mg.make_scene(name, depth)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    white = case.endswith("xorshift")
    ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_WHITE if white else rtb.abi.NOISE_PHILOX)
    try:
        simple = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
    finally:
        ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_PHILOX)
    assert np.array_equal(simple.out_color, g["color"]) and np.array_equal(simple.out_normal, g["normal"])
    assert np.array_equal(simple.out_albedo, g["albedo"]) and np.array_equal(simple.out_weight, g["weight"])
    assert np.array_equal(simple.diagnostics["ray_count"], g["ray_count"])
    if not white:
        mega = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        assert np.array_equal(mega.out_color[:, 3], g["color"][:, 3])
        assert np.array_equal(mega.diagnostics["ray_count"], g["ray_count"])
        n = np.maximum(g["color"][:, 3:4], 1)
        assert np.abs(mega.out_color[:, :3] / n - g["color"][:, :3] / n).max() <= RGB_TOL


@pytest.mark.parametrize("case", ["three_spheres_32x18x4_d8_philox", "final_linear_32x18x4_d50_philox",
                                  "final_bvh16_defocus_48x27x8_d50_philox", "mesh_bvh16_48x27x8_d50_philox"])
def test_gpu_matches_golden_fixtures(rtb, ctx, case):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name, depth, W, H, spp, td, ap, _ = mg.CASES[case]
    scene = mg.make_scene(name, depth)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    simple = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
    assert np.array_equal(simple.out_color, g["color"]) and np.array_equal(simple.out_normal, g["normal"])
    assert np.array_equal(simple.out_albedo, g["albedo"]) and np.array_equal(simple.out_weight, g["weight"])
    mega = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    assert np.array_equal(mega.out_color[:, 3], g["color"][:, 3])
    assert np.array_equal(mega.diagnostics["ray_count"], g["ray_count"])
    n = np.maximum(g["color"][:, 3:4], 1)
    assert np.abs(mega.out_color[:, :3] / n - g["color"][:, :3] / n).max() <= RGB_TOL


@pytest.mark.parametrize("name,depth,td,ap", [("final", 16, 50, 0.1), ("final", 32, 50, 0.0), ("three_spheres", 2, 50, 0.2)])
def test_image_does_not_depend_on_the_device_tree(rtb, oracle, ctx, name, depth, td, ap):
    """RTB_OPT_LEAF_SPHERES: collapsing subtrees of the host's BVH into device leaves changes what the
    walk executes (bounds_hit_count / candidate_count) and nothing else — every output is bitwise
    identical for every setting, and equal to the oracle's decisions."""
    W, H, spp = 96, 54, 24
    scene = rtb.host.make_scene(name, max_bvh_depth=depth)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    base = None
    try:
        for k in (1, 2, 5, 8, 15):
            ctx.set_option(rtb.abi.OPT_LEAF_SPHERES, k)
            got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
            assert_parity(ref, got, exact=False)
            simple = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
            assert_parity(ref, simple, exact=True)
            if base is None:
                base = got
            else:
                for x, y in ((base.out_color, got.out_color), (base.out_normal, got.out_normal), (base.out_albedo, got.out_albedo),
                             (base.out_weight, got.out_weight), (base.diagnostics["ray_count"], got.diagnostics["ray_count"])):
                    assert x.tobytes() == y.tobytes()
        # the exact re-test of the skipped host boxes, forced for every accepted hit: same image
        ctx.set_option(rtb.abi.OPT_LEAF_SPHERES, 8)
        ctx.set_option(rtb.abi.OPT_ALWAYS_WALK_CHAINS, 1)
        walked = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        assert walked.out_color.tobytes() == base.out_color.tobytes()
        assert walked.diagnostics["ray_count"].tobytes() == base.diagnostics["ray_count"].tobytes()
    finally:
        ctx.set_option(rtb.abi.OPT_LEAF_SPHERES, 1)
        ctx.set_option(rtb.abi.OPT_ALWAYS_WALK_CHAINS, 0)


@pytest.mark.parametrize("kernel", ["simple", "mega"])
def test_stress_scene_walks_the_world_in_hbm(rtb, oracle, ctx, kernel):
    """BASELINE config 5's world (10 000 dart-thrown spheres, BVH depth 16): the flattened world (~0.9 MB) does
    not fit shared memory, so the kernels read it in place through the read-only path — same decisions."""
    W, H, spp = 96, 54, 4
    scene = rtb.host.make_scene("stress", max_bvh_depth=16, target_count=10000)
    assert len(scene.spheres) == 10004
    assert rtb.plugin.describe_scene(scene, 1)["blob_bytes"] > 232448
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("depth,aperture", [(16, 0.0), (2, 0.15), (0, 0.0)])
def test_mesh_world_matches_the_oracle(rtb, oracle, ctx, kernel, depth, aperture):
    """rtb_upload_world: EntityType.Triangle entities (what the reference's host ingests at HEAD,
    Raytracer.cs:1185-1304) mixed with spheres — flat and smooth shading, two-sided hits, multi-entity leaves."""
    W, H, spp = 112, 63, 12
    scene = rtb.host.make_mesh_scene(max_bvh_depth=depth)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=aperture)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("depth,moving,aperture", [(16, True, 0.0), (2, True, 0.2), (0, False, 0.0), (16, False, 0.0)])
def test_cornell_world_of_placed_entities_matches_the_oracle(rtb, oracle, ctx, kernel, depth, moving, aperture):
    """rtb_upload_placed_world: EntityType.Rect / EntityType.Box entities behind rotated transforms, a sphere and a
    box that move during the exposure (Entity.TransformAtTime at Ray.Time), an emissive Rect as the only light
    (HitTests.cs:62-111, Entity.cs:57-127) — same decisions as the oracle for every path."""
    W, H, spp = 96, 54, 16
    scene = rtb.host.make_cornell_scene(max_bvh_depth=depth, moving=moving)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=aperture)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    assert ref.rgb().max() > 1.0 and ref.diagnostics["ray_count"].mean() > 2 * spp      # lit, and paths bounce
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


def test_cornell_world_white_noise_stream(rtb, oracle, ctx):
    """The reference's own xorshift32 stream (RTB_OPT_NOISE = 1): Ray.Time is the draw after the lens draws
    (View.cs:47) and the moving entities read it — bit-identical to the oracle in xorshift mode."""
    W, H, spp = 64, 36, 8
    scene = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref, noise=oracle.NOISE_XORSHIFT)
    ctx.upload(scene)
    ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_SIMPLE)
    ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_WHITE)
    try:
        got = rtb.plugin.HostBuffers(W, H)
        ctx.sample_batch(p, got)
    finally:
        ctx.set_option(rtb.abi.OPT_NOISE, 0)
    assert_parity(ref, got, exact=True)


def test_moving_entities_blur_and_static_ones_do_not(rtb, ctx):
    """Motion shows up only through Ray.Time: freezing the time range (the whole motion happens before t = 0) moves
    the entities to their destination for every ray; the default range spreads them over the exposure."""
    W, H, spp = 96, 54, 32
    scene = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True)
    p = rtb.host.make_params(scene, W, H, spp, 50)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    frozen = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True)
    frozen.placed["time_range"][frozen.placed["moving"] == 1] = (-2.0, -1.0)
    b = render_gpu(rtb, ctx, frozen, p, W, H, rtb.abi.KERNEL_MEGA)
    still = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True)
    m = still.placed["moving"] == 1
    still.placed["position"][m] += still.placed["destination_offset"][m]
    still.placed["moving"][m] = 0
    c = render_gpu(rtb, ctx, still, p, W, H, rtb.abi.KERNEL_MEGA)
    na, nb, nc = (x.out_normal / np.maximum(x.out_color[:, 3:4], 1) for x in (a, b, c))
    assert np.abs(na - nb).max() > 0.2            # the blurred silhouette differs from the frozen one
    assert np.abs(nb - nc).max() <= 1e-3          # frozen at the destination == a static entity placed there


@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("depth,emissive", [(16, False), (16, True), (2, False)])
def test_textured_mesh_world_matches_the_oracle(rtb, oracle, ctx, kernel, depth, emissive):
    """rtb_upload_textures: TextureType.Image albedo / emission / glossiness (alpha of an RGBA32 map) / metallic maps on mesh
    materials, sampled at the triangles' interpolated vertex uvs (Texture.cs:80-89,128-137; HitTests.cs:147); sphere entities
    sample texel (0, 0) (Entity.cs:107) — including a Dielectric whose roughness comes from a map."""
    W, H, spp = 112, 63, 12
    scene = rtb.host.make_mesh_scene(max_bvh_depth=depth, emissive=emissive, textured=True)
    p = rtb.host.make_params(scene, W, H, spp, 12 if emissive else 50)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    plain = oracle.Buffers(W, H)
    oracle.sample_batch(rtb.host.make_mesh_scene(max_bvh_depth=depth, emissive=emissive), p, plain)
    assert np.abs(ref.rgb() - plain.rgb()).max() > 0.2                       # the maps are visible
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))
    # a later world upload drops the textures
    got2 = render_gpu(rtb, ctx, rtb.host.make_mesh_scene(max_bvh_depth=depth, emissive=emissive), p, W, H, k)
    assert_parity(plain, got2, exact=(kernel == "simple"))


@pytest.mark.parametrize("kernel", ["simple", "mega"])
def test_textured_cornell_world_matches_the_oracle(rtb, oracle, ctx, kernel):
    """Placed entities with image-textured materials (the kernel flavour that carries both): Rect / Box / sphere entities
    have TexCoords = 0, so each takes texel (0, 0) of its maps — albedo, a glossiness map on the metal box (no longer a
    perfect mirror: Material.cs:190-192) and an emission map on the light."""
    W, H, spp = 96, 54, 16
    scene = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True)
    mt = np.zeros(len(scene.materials), dtype=rtb.abi.MATERIAL_TEXTURES_DTYPE)
    for k in ("albedo_image", "emission_image", "glossiness_image", "metallic_image"):
        mt[k] = -1
    mt[0]["albedo_image"] = 1
    mt[3]["emission_image"] = 3
    mt[5]["glossiness_image"], mt[5]["glossiness_channel"] = 1, 3
    mt[6]["albedo_image"], mt[6]["metallic_image"] = 0, 2
    scene.images, scene.material_textures, scene.triangle_uvs = rtb.host._test_images(), mt, None
    p = rtb.host.make_params(scene, W, H, spp, 50)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    plain = oracle.Buffers(W, H)
    oracle.sample_batch(rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True), p, plain)
    assert np.abs(ref.rgb() - plain.rgb()).max() > 0.2
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


def test_textures_on_a_sphere_world_and_argument_checks(rtb, oracle, ctx):
    """Sphere entities have TexCoords = 0: a textured material on a sphere world takes texel (0, 0) everywhere — and moves the
    world to the general kernel flavour.  Wrong counts / channels are rejected."""
    W, H, spp = 64, 36, 8
    scene = rtb.host.make_scene("three_spheres", max_bvh_depth=2)
    img = np.zeros((4, 4, 3), np.uint8)
    img[0, 0] = (255, 128, 0)
    img[1:, 1:] = (0, 0, 255)
    mt = np.zeros(len(scene.materials), dtype=rtb.abi.MATERIAL_TEXTURES_DTYPE)
    for k in ("albedo_image", "emission_image", "glossiness_image", "metallic_image"):
        mt[k] = -1
    mt[0]["albedo_image"] = 0
    scene.images, scene.material_textures, scene.triangle_uvs = [img], mt, None
    p = rtb.host.make_params(scene, W, H, spp, 50)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    for k in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
        got = render_gpu(rtb, ctx, scene, p, W, H, k)
        assert_parity(ref, got, exact=(k == rtb.abi.KERNEL_SIMPLE))
    with pytest.raises(rtb.plugin.RtbError):
        ctx.upload_textures([img], mt[:-1])                                    # material count differs from the world's
    bad = mt.copy()
    bad[1]["glossiness_image"], bad[1]["glossiness_channel"] = 0, 3            # alpha of an RGB24 image
    with pytest.raises(rtb.plugin.RtbError):
        ctx.upload_textures([img], bad)
    with pytest.raises(rtb.plugin.RtbError):
        ctx.set_option(rtb.abi.OPT_KERNEL, 3)                                  # the rejected wavefront kernel is gone
    ctx.set_option(rtb.abi.OPT_KERNEL, 0)


@pytest.mark.parametrize("kernel", ["simple", "mega"])
def test_emissive_panel_without_sky(rtb, oracle, ctx, kernel):
    """Material.Emit (Material.cs:175-179) + SkyType.None: the overhead panel is the only light, so every pixel's
    radiance is emission carried down the path by the attenuation product (SampleBatchJob.cs:383-396)."""
    W, H, spp = 96, 54, 16
    scene = rtb.host.make_mesh_scene(max_bvh_depth=16, emissive=True)
    p = rtb.host.make_params(scene, W, H, spp, 12)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    assert ref.rgb().max() > 0.5 and (ref.rgb().reshape(-1, 3).max(axis=1) == 0).mean() > 0.05   # lit floor, black sky
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


@pytest.mark.parametrize("kernel", ["simple", "mega"])
def test_emissive_sphere_in_a_sphere_world(rtb, oracle, ctx, kernel):
    """Material.Emit on sphere entities (the lean sphere build carries radiance along the path too)."""
    W, H, spp = 80, 45, 16
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    scene.materials["emission"][scene.spheres["material"][5]] = (4.0, 3.0, 2.0)
    scene.materials["emission"][scene.spheres["material"][40]] = (0.0, 2.5, 6.0)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    dark = rtb.host.make_scene("final", max_bvh_depth=16)
    base = oracle.Buffers(W, H)
    oracle.sample_batch(dark, p, base)
    assert np.abs(ref.rgb() - base.rgb()).max() > 0.5            # the emitters are visible
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    got = render_gpu(rtb, ctx, scene, p, W, H, k)
    assert_parity(ref, got, exact=(kernel == "simple"))


def _test_cubemap(size=16):
    """A deterministic HDR sky: a different tint per face, a gradient across it and one bright 'sun' texel."""
    rng = np.random.default_rng(11)
    faces = np.zeros((6, size, size, 4), np.float32)
    tint = np.array([(1, .8, .6), (.5, .7, 1), (.9, .95, 1), (.3, .25, .2), (.7, 1, .7), (1, .6, .9)], np.float32)
    g = np.linspace(0.4, 1.2, size, dtype=np.float32)
    for f in range(6):
        faces[f, :, :, :3] = tint[f] * g[None, :, None] * g[::-1][:, None, None] + 0.05 * rng.random((size, size, 3), dtype=np.float32)
    faces[2, size // 3, size // 2, :3] = (40.0, 36.0, 30.0)
    faces[..., 3] = 1
    return faces.astype(np.float16)


@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("world", ["mesh", "final"])
def test_cubemap_sky(rtb, oracle, ctx, kernel, world):
    """SkyType.CubeMap (Environment.cs, Texture.cs:141-211) — the sky the reference's host builds from the scene's HDRI
    sky at HEAD (Raytracer.cs:663-665): nearest-texel lookups of R16G16B16A16_SFloat faces."""
    W, H, spp = 96, 54, 12
    scene = rtb.host.make_mesh_scene(max_bvh_depth=16) if world == "mesh" else rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.05)
    p.environment.sky_type = rtb.abi.SKY_CUBEMAP
    faces = _test_cubemap()
    oracle.set_sky_cubemap(faces)
    ctx.upload_sky_cubemap(faces)
    try:
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, p, ref)
        assert ref.rgb().max() > 2.0                            # the sun texel is seen (directly or in a reflection)
        k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
        got = render_gpu(rtb, ctx, scene, p, W, H, k)
        assert_parity(ref, got, exact=(kernel == "simple"))
    finally:
        oracle.set_sky_cubemap(None)
        ctx.upload_sky_cubemap(None)
    with pytest.raises(rtb.plugin.RtbError) as e:               # cube-map sky requested, none uploaded
        ctx.sample_batch(p, rtb.plugin.HostBuffers(W, H))
    assert e.value.code == rtb.abi.RTB_ERR_NO_SCENE


@pytest.mark.parametrize("world,depth,td,ap", [("three_spheres", 0, 8, None), ("final", 16, 50, 0.1), ("final", 3, 12, 0.0), ("mesh", 16, 50, 0.0)])
def test_reference_white_noise_stream_is_reproduced_bit_for_bit(rtb, oracle, ctx, world, depth, td, ap):
    """RTB_OPT_NOISE = 1: the thread-per-pixel kernel consumes the reference's OWN random stream (NoiseColor.White:
    Unity.Mathematics.Random seeded (Seed * 0x8C4CA03F) ^ (index * 0x7383ED49), SampleBatchJob.cs:91, draws made and
    skipped exactly as Material.Scatter / View.GetRay make them) and equals the CPU restatement run with that
    generator in every output bit — the algorithm AND the generator of the reference, not only the Philox substitute."""
    W, H, spp = 72, 40, 6
    scene = rtb.host.make_mesh_scene(max_bvh_depth=depth) if world == "mesh" else rtb.host.make_scene(world, max_bvh_depth=depth)
    p = rtb.host.make_params(scene, W, H, spp, td, aperture=ap, seed=3)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref, noise=oracle.NOISE_XORSHIFT)
    philox = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, philox, noise=oracle.NOISE_PHILOX)
    assert not np.array_equal(ref.out_color, philox.out_color)          # a different stream, a different image
    ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_WHITE)
    try:
        got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
        assert_parity(ref, got, exact=True)
        with pytest.raises(rtb.plugin.RtbError) as e:                   # a sequential stream cannot be split over lanes
            render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        assert e.value.code == rtb.abi.RTB_ERR_UNSUPPORTED
    finally:
        ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_PHILOX)
        ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_AUTO)
    if world == "three_spheres":                                         # and the committed xorshift fixture
        g = np.load(os.path.join(GOLDEN, "three_spheres_32x18x4_d8_xorshift.npz"))
        p2 = rtb.host.make_params(scene, 32, 18, 4, 8)
        ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_WHITE)
        try:
            fx = render_gpu(rtb, ctx, scene, p2, 32, 18, rtb.abi.KERNEL_SIMPLE)
        finally:
            ctx.set_option(rtb.abi.OPT_NOISE, rtb.abi.NOISE_PHILOX)
        assert np.array_equal(fx.out_color, g["color"]) and np.array_equal(fx.out_normal, g["normal"])
        assert np.array_equal(fx.out_albedo, g["albedo"]) and np.array_equal(fx.out_weight, g["weight"])


def test_world_upload_rejects_what_it_cannot_render(rtb, ctx):
    scene = rtb.host.make_mesh_scene()
    ents = scene.entities.copy()
    ents["type"][0] = 9                                      # not an EntityType
    with pytest.raises(rtb.plugin.RtbError) as e:
        ctx.upload_world(ents, scene.spheres, scene.triangles, scene.materials, scene.nodes)
    assert e.value.code == rtb.abi.RTB_ERR_UNSUPPORTED
    ents = scene.entities.copy()
    ents["type"][0] = rtb.abi.ENTITY_BOX                     # a Box is a placed entity: there is no placed array in this call
    with pytest.raises(rtb.plugin.RtbError) as e:
        ctx.upload_world(ents, scene.spheres, scene.triangles, scene.materials, scene.nodes)
    assert e.value.code == rtb.abi.RTB_ERR_INVALID_ARGUMENT
    ents = scene.entities.copy()
    ents["index"][0] = 10 ** 6
    with pytest.raises(rtb.plugin.RtbError) as e:
        ctx.upload_world(ents, scene.spheres, scene.triangles, scene.materials, scene.nodes)
    assert e.value.code == rtb.abi.RTB_ERR_INVALID_ARGUMENT


def test_pinned_host_arrays_are_used_in_place(rtb, ctx):
    """rtb_sample_batch on registered (pinned) host arrays runs the kernel on them in place; pageable arrays are
    staged through device copies.  Same bytes either way, also with interlaced rows and a previous accumulation."""
    W, H, spp = 160, 90, 8
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    ctx.upload(scene)
    ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_MEGA)
    rng = np.random.default_rng(5)
    for divider, offset in ((1, 0), (3, 1)):
        p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, slice_offset=offset, slice_divider=divider)
        staged, pinned = rtb.plugin.HostBuffers(W, H), rtb.plugin.HostBuffers(W, H)
        prev = rng.random((W * H, 4)).astype(np.float32)
        prev[:, 3] = 3
        for b in (staged, pinned):
            b.in_color[:] = prev
            b.in_weight[:] = 1.5
            b.out_color[:] = -7.0              # rows the batch skips must keep this
        ctx.sample_batch(p, staged)
        assert not ctx.last_batch_in_place()
        ctx.register_host_buffers(pinned)
        try:
            ctx.sample_batch(p, pinned)
            assert ctx.last_batch_in_place()
            ctx.set_option(rtb.abi.OPT_HOST_ACCESS, 0)
            again = rtb.plugin.HostBuffers(W, H)
            again.in_color[:] = prev
            again.in_weight[:] = 1.5
            again.out_color[:] = -7.0
            ctx.sample_batch(p, again)
            assert not ctx.last_batch_in_place()
        finally:
            ctx.set_option(rtb.abi.OPT_HOST_ACCESS, 1)
            ctx.unregister_host_buffers(pinned)
        for x, y in zip(staged.arrays()[4:], pinned.arrays()[4:]):
            assert x.tobytes() == y.tobytes()
        assert staged.out_color.tobytes() == again.out_color.tobytes()


def test_interlaced_rows_and_carry_over(rtb, oracle, ctx):
    """SliceOffset/SliceDivider (SampleBatchJob.cs:69): skipped rows keep whatever the host put in out_*."""
    W, H, spp = 48, 30, 4
    scene = rtb.host.make_scene("three_spheres")
    for off in (0, 2):
        p = rtb.host.make_params(scene, W, H, spp, 8, slice_offset=off, slice_divider=3)
        ref = oracle.Buffers(W, H)
        ref.out_color[:] = 7.0
        oracle.sample_batch(scene, p, ref)
        for k in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
            ctx.upload(scene)
            ctx.set_option(rtb.abi.OPT_KERNEL, k)
            got = rtb.plugin.HostBuffers(W, H)
            got.out_color[:] = 7.0
            ctx.sample_batch(p, got)
            rows = np.arange(H)
            active = np.repeat(rows % 3 == off, W)
            assert np.all(got.out_color[~active] == 7.0)
            assert np.array_equal(ref.out_color[:, 3], got.out_color[:, 3])
            assert np.abs(ref.rgb() - got.rgb())[active.reshape(H, W)].max() <= RGB_TOL


def test_row_tiles_compose_to_the_full_frame_bitwise(rtb, ctx):
    """Row-tile sharding (rtb extension row_begin/row_end): tiles rendered separately are bit-identical
    to the whole frame because Philox is keyed by the global pixel index."""
    W, H, spp = 96, 54, 8
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    full = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    tiled = rtb.plugin.HostBuffers(W, H)
    for b, e in [(0, 7), (7, 20), (20, 54)]:
        pt = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, row_begin=b, row_end=e)
        ctx.sample_batch(pt, tiled)
    for a, b_ in zip(full.arrays(), tiled.arrays()):
        assert a.tobytes() == b_.tobytes()        # bytes: sample_count_weight diagnostics are NaN on a fresh buffer


def test_progressive_accumulation_two_batches(rtb, oracle, ctx):
    """accumulators in -> accumulators out across batches (Raytracer.cs:798-802), seeds 1 and 2."""
    W, H, spp = 64, 36, 6
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    ref, ctxb = oracle.Buffers(W, H), {}
    for k in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
        ctxb[k] = rtb.plugin.HostBuffers(W, H)
    ctx.upload(scene)
    for batch in range(2):
        p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, seed=1 + batch)
        if batch:
            ref.swap()
        oracle.sample_batch(scene, p, ref)
        for k, b in ctxb.items():
            if batch:
                b.swap()
            ctx.set_option(rtb.abi.OPT_KERNEL, k)
            ctx.sample_batch(p, b)
    assert ref.out_color[:, 3].max() == 2 * spp
    assert np.array_equal(ref.out_color, ctxb[rtb.abi.KERNEL_SIMPLE].out_color)
    assert np.array_equal(ref.out_weight, ctxb[rtb.abi.KERNEL_SIMPLE].out_weight)
    assert_parity(ref, ctxb[rtb.abi.KERNEL_MEGA], exact=False)


def test_adaptive_sample_counts(rtb, oracle, ctx):
    """SampleCountRange.x < y with a fed-back weight (SampleBatchJob.cs:118-126): ragged per-pixel work."""
    W, H = 48, 27
    scene = rtb.host.make_scene("three_spheres")
    p0 = rtb.host.make_params(scene, W, H, 4, 8)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p0, ref)
    sc = ref.out_color[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        w = ref.out_weight / sc
    ext = (float(np.nanmin(w)), float(np.nanmax(w)))
    ref.swap()
    p1 = rtb.host.make_params(scene, W, H, 2, 8, seed=2, spp_max=9)
    p1.sample_count_weight_extrema[0], p1.sample_count_weight_extrema[1] = ext
    inputs = (ref.in_color.copy(), ref.in_weight.copy(), ref.in_normal.copy(), ref.in_albedo.copy())
    oracle.sample_batch(scene, p1, ref)
    added = ref.diagnostics["ray_count"]
    assert added.min() >= 2 and len(np.unique(ref.out_color[:, 3] - inputs[0][:, 3])) > 2   # truly ragged
    for k, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
        got = render_gpu(rtb, ctx, scene, p1, W, H, k, inputs=inputs)
        assert_parity(ref, got, exact)


def test_all_samples_fail_fallback_outputs(rtb, oracle, ctx):
    """TraceDepth 1: paths that hit anything fail; count 0 pixels emit the first sample's normal/albedo
    (SampleBatchJob.cs:152-161)."""
    W, H = 40, 24
    scene = rtb.host.make_scene("three_spheres")
    p = rtb.host.make_params(scene, W, H, 3, 1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    assert (ref.out_color[:, 3] == 0).any()
    for k in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
        got = render_gpu(rtb, ctx, scene, p, W, H, k)
        zero = ref.out_color[:, 3] == 0
        assert np.array_equal(ref.out_color[:, 3], got.out_color[:, 3])
        assert np.array_equal(ref.out_normal[zero], got.out_normal[zero])
        assert np.array_equal(ref.out_albedo[zero], got.out_albedo[zero])


def test_empty_world_and_single_sphere(rtb, oracle, ctx):
    W, H = 32, 18
    scene = rtb.host.make_scene("three_spheres")
    p = rtb.host.make_params(scene, W, H, 2, 8)
    empty_nodes = np.zeros(0, rtb.abi.BVH_NODE_DTYPE)
    for spheres, nodes in [(scene.spheres[:0], empty_nodes), rtb.host.build_bvh(scene.spheres[2:3], 0)]:
        class S:  # a scene-like holder for the oracle wrapper
            pass
        s = S()
        s.spheres, s.materials, s.nodes = np.ascontiguousarray(spheres), scene.materials, nodes
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(s, p, ref)
        for k, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
            ctx.upload_scene(s.spheres, s.materials, s.nodes)
            ctx.set_option(rtb.abi.OPT_KERNEL, k)
            got = rtb.plugin.HostBuffers(W, H)
            ctx.sample_batch(p, got)
            assert_parity(ref, got, exact)


def test_determinism_and_counters(rtb, ctx):
    W, H, spp = 160, 90, 16
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    b = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    assert a.out_color.tobytes() == b.out_color.tobytes() and a.out_weight.tobytes() == b.out_weight.tobytes()
    ctx.set_option(rtb.abi.OPT_COUNTERS, 1)
    try:
        c = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        cnt = ctx.counters()
    finally:
        ctx.set_option(rtb.abi.OPT_COUNTERS, 0)
    assert c.out_color.tobytes() == a.out_color.tobytes()
    assert cnt["samples"] == W * H * spp
    assert cnt["rays"] == int(a.diagnostics["ray_count"].astype(np.int64).sum())
    assert cnt["samples"] - cnt["failed_samples"] == int(a.out_color[:, 3].astype(np.int64).sum()) == cnt["sky_hits"]
    assert cnt["shade_standard"] + cnt["shade_dielectric"] + cnt["sky_hits"] == cnt["rays"]
    assert cnt["node_tests"] > cnt["rays"] and cnt["sphere_tests"] > 0


def test_error_behaviour(rtb):
    abi = rtb.abi
    c = rtb.plugin.Context(0)
    try:
        scene = rtb.host.make_scene("three_spheres")
        p = rtb.host.make_params(scene, 16, 9, 1, 4)
        b = rtb.plugin.HostBuffers(16, 9)
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, b)
        assert e.value.code == abi.RTB_ERR_NO_SCENE
        bad = scene.materials.copy()
        bad["type"][0] = 7                                # not a MaterialType
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.upload_scene(scene.spheres, bad, scene.nodes)
        assert e.value.code == abi.RTB_ERR_UNSUPPORTED
        nodes = scene.nodes.copy()
        nodes["entity_count"][0] = 99
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.upload_scene(scene.spheres, scene.materials, nodes)
        assert e.value.code == abi.RTB_ERR_INVALID_ARGUMENT
        c.upload(scene)
        p.slice_divider = 0
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, b)
        assert e.value.code == abi.RTB_ERR_INVALID_ARGUMENT
        p.slice_divider = 1
        p.environment.sky_type = abi.SKY_CUBEMAP          # cube-map sky without rtb_upload_sky_cubemap
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, b)
        assert e.value.code == abi.RTB_ERR_NO_SCENE
        p.environment.sky_type = 7
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, b)
        assert e.value.code == abi.RTB_ERR_INVALID_ARGUMENT
        p.environment.sky_type = abi.SKY_GRADIENT
        cancel = np.ones(1, np.uint8)
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, b, cancel=cancel)
        assert e.value.code == abi.RTB_ERR_CANCELLED
        cancel[0] = 0
        c.sample_batch(p, b, cancel=cancel)          # a live token that is never set: the same single launch, the same image
        b2 = rtb.plugin.HostBuffers(16, 9)
        c.sample_batch(p, b2)
        assert b.out_color.tobytes() == b2.out_color.tobytes()
    finally:
        c.close()


def test_combine_and_reduce_metrics_device(rtb, ctx):
    """CombineJob / ReduceMetricsJob on device buffers against their numpy restatement."""
    import torch

    W, H, spp = 64, 36, 4
    scene = rtb.host.make_scene("three_spheres")
    p = rtb.host.make_params(scene, W, H, spp, 8, slice_offset=1, slice_divider=2)   # half the rows stay empty
    ctx.upload(scene)
    ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_MEGA)
    n = W * H
    dev = torch.device("cuda:0")
    z = lambda c: torch.zeros(n, c, device=dev, dtype=torch.float32)   # noqa: E731
    in_c, in_w, in_n, in_a = z(4), torch.zeros(n, device=dev), z(3), z(3)
    out_c, out_w, out_n, out_a = z(4), torch.zeros(n, device=dev), z(3), z(3)
    diag = torch.zeros(n, 4, device=dev)
    bufs = rtb.plugin.device_buffers_struct(in_c, in_w, in_n, in_a, out_c, out_w, out_n, out_a, diag)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.sample_batch_device(p, bufs, stream)
    fc, fn, fa = z(3), z(3), z(3)
    ctx.combine_device(W, H, out_c, out_n, out_a, fc, fn, fa, stream=stream)
    m = ctx.reduce_metrics_device(W, H, diag, out_c, out_w, stream=stream)
    torch.cuda.synchronize()
    c = out_c.cpu().numpy()
    cnt = c[:, 3].astype(np.int32).reshape(H, W)
    # depth 8: paths caught inside the hollow glass sphere fail, so active rows hold 0..spp samples
    assert cnt[1::2].max() == spp and np.median(cnt[1::2]) == spp and (cnt[0::2] == 0).all()
    col = c[:, :3].reshape(H, W, 3)
    want = np.zeros((H, W, 3), np.float32)            # CombineJob restated: look down the column until a sampled pixel
    for y in range(H):
        for x in range(W):
            yy = y
            while cnt[yy, x] == 0 and yy - 1 >= 0:
                yy -= 1
            if cnt[yy, x] > 0:
                want[y, x] = col[yy, x] / np.float32(cnt[yy, x])
    assert np.abs(fc.cpu().numpy().reshape(H, W, 3) - want).max() < 1e-6
    nn = fn.cpu().numpy()
    norms = np.linalg.norm(nn, axis=1).reshape(H, W)
    assert np.allclose(norms[1::2][cnt[1::2] > 0], 1, atol=1e-5) and (norms[0::2] == 0).all()   # normalizesafe
    assert m.total_samples == int(cnt.sum()) and m.sample_count_min == 0 and m.sample_count_max == spp
    assert m.total_ray_count == int(diag[:, 0].sum().item())
    with np.errstate(divide="ignore", invalid="ignore"):
        w = out_w.cpu().numpy().reshape(H, W) / cnt.astype(np.float32)
    assert abs(m.sample_count_weight_min - np.nanmin(w)) < 1e-6 and abs(m.sample_count_weight_max - np.nanmax(w)) < 1e-6


def test_full_size_properties_config3(rtb, oracle, ctx):
    """BASELINE config 3 at full size (1920x1080, 256 spp, depth 50, BVH + defocus): properties that do
    not need the oracle — every sample accounted for, determinism, shard-invariance on a band, and the
    downsampled image agreeing with an independent 256-spp oracle render statistically."""
    W, H, spp = 1920, 1080, 256
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    cnt = a.out_color[:, 3]
    # failed samples (depth == TraceDepth, SampleBatchJob.cs:379-381) are rare glass paths
    assert cnt.max() == spp and cnt.min() >= spp // 2 and (W * H * spp - cnt.astype(np.int64).sum()) < 1e-3 * W * H * spp
    # the row holding the pixel with the most failed samples, against the oracle: counts and ray counts exact
    worst = int(np.argmin(cnt)) // W
    pr = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, row_begin=worst, row_end=worst + 1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, pr, ref)
    rs = slice(worst * W, (worst + 1) * W)
    assert np.array_equal(ref.out_color[rs, 3], a.out_color[rs, 3])
    assert np.array_equal(ref.diagnostics["ray_count"][rs], a.diagnostics["ray_count"][rs])
    assert np.abs(ref.rgb()[worst] - a.rgb()[worst]).max() <= RGB_TOL
    assert np.isfinite(a.out_color).all()
    rays = a.diagnostics["ray_count"].astype(np.int64)
    assert rays.min() >= spp and rays.sum() > 2 * W * H * spp
    rgb = a.rgb()
    assert 0.0 <= rgb.min() and rgb.max() <= 1.0 + 1e-5          # no emitters: radiance bounded by the sky
    # a band of rows rendered on its own is bit-identical to the same rows of the full frame
    band = rtb.plugin.HostBuffers(W, H)
    pb = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, row_begin=500, row_end=516)
    ctx.sample_batch(pb, band)
    sl = slice(500 * W, 516 * W)
    assert band.out_color[sl].tobytes() == a.out_color[sl].tobytes()
    assert band.out_normal[sl].tobytes() == a.out_normal[sl].tobytes()
    # sky gradient at the top rows (no geometry there): exact analytic check of the miss path
    top = rgb[H - 4:, :, :]
    assert np.abs(top[..., 0] - top[..., 0].mean()).max() < 0.02 and top[..., 2].min() > 0.95


def test_finalize_device_matches_the_restated_job_and_runs_at_hbm_rate(rtb, oracle, ctx):
    """FinalizeTexturesJob (FinalizeTexturesJob.cs:23-55) on device buffers: bytes equal the numpy restatement that
    uses the oracle's pow; at 4K the kernel is HBM-bound (36 B read + 12 B written per pixel)."""
    import torch

    dev = torch.device("cuda:0")
    W, H = 3840, 2160
    n = W * H
    g = torch.Generator(device="cpu").manual_seed(3)
    col = (torch.rand(n, 3, generator=g) * 1.6 - 0.2)
    col[:7] = torch.tensor([[0, 0, 0], [1, 1, 1], [2, 0.5, -1], [1e-8, 0.0031308, 0.5], [0.25, 0.75, 0.999], [float("nan"), 0.5, 0.5], [1e9, 1e-30, 0.2]])
    nor = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    alb = torch.rand(n, 3, generator=g)
    d_col, d_nor, d_alb = col.to(dev), nor.to(dev), alb.to(dev)
    o = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(3)]
    stream = torch.cuda.current_stream().cuda_stream
    ctx.finalize_device(W, H, d_col, d_nor, d_alb, o[0], o[1], o[2], stream=stream)
    torch.cuda.synchronize()

    def want(x):                          # LinearToGamma + saturate * 255 -> byte, restated with the shared pow
        x = np.maximum(x.astype(np.float32), np.float32(0))       # math.max(v, 0): NaN -> 0 ("isnan(y) || x > y ? x : y" with x = v)
        flat = np.ascontiguousarray(x.reshape(-1))
        p = np.zeros_like(flat)
        pos = flat > 0
        src = np.ascontiguousarray(flat[pos])
        dst = np.zeros_like(src)
        oracle.lib().oracle_umath_pow(src.ctypes.data, np.float32(0.416666667), dst.ctypes.data, len(src))
        p[pos] = dst
        gmm = np.maximum(np.float32(1.055) * p - np.float32(0.055), np.float32(0))
        b = (np.clip(gmm, 0, 1).astype(np.float32) * np.float32(255)).astype(np.uint32).reshape(-1, 3)
        return b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | np.uint32(0xff000000)

    sample = slice(0, 200000)
    got = [t.cpu().numpy().view(np.uint32) for t in o]
    keep = np.r_[0:5, 6:sample.stop]                              # row 5 holds a NaN (checked below)
    assert np.array_equal(got[0][keep], want(col.numpy()[keep]))
    assert np.array_equal(got[1][sample], want(nor.numpy()[sample] * np.float32(0.5) + np.float32(0.5)))
    assert np.array_equal(got[2][sample], want(alb.numpy()[sample]))
    # black stays 0; white is 254, not 255: 1.055f - 0.055f = 0.99999994 in float, truncated after * 255 (the job's arithmetic)
    assert got[0][0] == 0xff000000 and got[0][1] == 0xfffefefe and (got[0][2] & 0xff) == 255 and (got[0][2] >> 16) & 0xff == 0
    assert (got[0][5] >> 8) & 0xff == (got[0][5] >> 16) & 0xff and got[0][5] >> 24 == 0xff       # NaN channel: any byte, no fault
    # bandwidth: CUDA events over 20 launches on inputs (3 x 99.5 MB) larger than L2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ctx.finalize_device(W, H, d_col, d_nor, d_alb, o[0], o[1], o[2], stream=stream)
    ev[0].record()
    for _ in range(20):
        ctx.finalize_device(W, H, d_col, d_nor, d_alb, o[0], o[1], o[2], stream=stream)
    ev[1].record()
    torch.cuda.synchronize()
    gbs = 20 * n * 48 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
    print(f"finalize_kernel: {gbs:.0f} GB/s algorithmic")
    assert gbs > 4000


def test_finalize_fast_path_equals_the_exact_pow_on_a_dense_sweep(rtb, oracle, ctx):
    """The finalize kernel takes a MUFU estimate of pow wherever it lies further from a byte boundary than it can be wrong, and
    the polynomial pow of the job's restatement elsewhere (aux_kernels.cuh: gamma_byte).  Every float k / 2^24 of [0, 1] — steps
    far finer than the guard, on both sides of all 255 boundaries — and a log sweep of [1e-7, 40] must give the restated job's
    byte; the ragged size also takes the kernel's per-pixel tail."""
    import torch

    dev = torch.device("cuda:0")
    dense = (np.arange(3 * 5592406, dtype=np.float64) / 2.0**24).astype(np.float32)
    logs = np.exp(np.linspace(np.log(1e-7), np.log(40.0), 3 * 700001)).astype(np.float32)
    for vals in (dense, logs):
        n = len(vals) // 3
        x = torch.from_numpy(vals.reshape(n, 3).copy()).to(dev)
        out = torch.zeros(n, dtype=torch.int32, device=dev)
        ctx.finalize_device(1, n, x, None, None, out, None, None, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint32)
        p = np.zeros_like(vals)
        pos = vals > 0
        src = np.ascontiguousarray(vals[pos])
        dst = np.zeros_like(src)
        oracle.lib().oracle_umath_pow(src.ctypes.data, np.float32(0.416666667), dst.ctypes.data, len(src))
        p[pos] = dst
        g = np.maximum(np.float32(1.055) * p - np.float32(0.055), np.float32(0))
        b = (np.clip(g, 0, 1).astype(np.float32) * np.float32(255)).astype(np.uint32).reshape(n, 3)
        want = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | np.uint32(0xff000000)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (len(bad), vals.reshape(n, 3)[bad[:5]], got[bad[:5]], want[bad[:5]])


def test_full_size_config2_rows_match_the_oracle(rtb, oracle, ctx):
    """BASELINE config 2 at full size (book-1 final scene as ONE leaf of 482 spheres — the linear hit list —, 1280x720, 64 spp,
    depth 50): every sample accounted for, and three whole rows (ground, sphere band, sky) equal the oracle's decisions exactly,
    RGB within tolerance.  This is the build with the deferred division in the big leaf (kFlavorChains)."""
    W, H, spp = 1280, 720, 64
    scene = rtb.host.make_scene("final", max_bvh_depth=0)
    p = rtb.host.make_params(scene, W, H, spp, 50)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    cnt = a.out_color[:, 3]
    assert cnt.max() == spp and (W * H * spp - cnt.astype(np.int64).sum()) < 1e-3 * W * H * spp
    assert np.isfinite(a.out_color).all()
    rgb = a.rgb()
    for row in (60, 300, 650):
        pr = rtb.host.make_params(scene, W, H, spp, 50, row_begin=row, row_end=row + 1)
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, pr, ref)
        rs = slice(row * W, (row + 1) * W)
        assert np.array_equal(ref.out_color[rs, 3], a.out_color[rs, 3])
        assert np.array_equal(ref.diagnostics["ray_count"][rs], a.diagnostics["ray_count"][rs])
        assert np.abs(ref.rgb()[row] - rgb[row]).max() <= RGB_TOL


def test_full_size_properties_config4(rtb, oracle, ctx):
    """BASELINE config 4 at full size on one GPU (3840x2160, 1024 spp, depth 50: 8.5 G camera paths): every sample
    accounted for, finite, radiance bounded by the sky, and one row checked against the oracle exactly."""
    W, H, spp = 3840, 2160, 1024
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    cnt = a.out_color[:, 3]
    assert cnt.max() == spp and cnt.min() >= spp // 2 and (W * H * spp - cnt.astype(np.int64).sum()) < 1e-3 * W * H * spp
    assert np.isfinite(a.out_color).all()
    rgb = a.rgb()
    assert 0.0 <= rgb.min() and rgb.max() <= 1.0 + 1e-5
    rays = a.diagnostics["ray_count"].astype(np.int64)
    assert rays.min() >= spp and rays.sum() > 2 * W * H * spp
    row = 700
    pr = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, row_begin=row, row_end=row + 1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, pr, ref)
    rs = slice(row * W, (row + 1) * W)
    assert np.array_equal(ref.out_color[rs, 3], a.out_color[rs, 3])
    assert np.array_equal(ref.diagnostics["ray_count"][rs], a.diagnostics["ray_count"][rs])
    assert np.abs(ref.rgb()[row] - rgb[row]).max() <= RGB_TOL


def test_full_size_properties_config5(rtb, oracle, ctx):
    """BASELINE config 5 at full size on one GPU (10 004-sphere stress world read in place from HBM, 4096x4096, 2048 spp,
    depth 50: 34 G camera paths, ~10 s): every sample accounted for, finite, bounded by the sky, the 2048 samples of a pixel
    survive the packed per-lane counters, and part of one row equals the oracle's decisions exactly."""
    W, H, spp = 4096, 4096, 2048
    scene = rtb.host.make_scene("stress", max_bvh_depth=16, target_count=10000)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    cnt = a.out_color[:, 3]
    # paths that are still bouncing between the glass spheres at depth 50 fail (0.44 % of this world's paths)
    assert cnt.max() == spp and cnt.min() >= spp // 2 and (W * H * spp - cnt.astype(np.int64).sum()) < 1e-2 * W * H * spp
    assert np.isfinite(a.out_color).all()
    rgb = a.rgb()
    assert 0.0 <= rgb.min() and rgb.max() <= 1.0 + 1e-5
    rays = a.diagnostics["ray_count"].astype(np.int64)
    assert rays.min() >= spp and rays.sum() > 1.5 * W * H * spp
    row, cols = 1500, 512                                # 512 pixels x 2048 spp through the oracle
    pr = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, pr, ref, index_range=(row * W + 1024, row * W + 1024 + cols))
    rs = slice(row * W + 1024, row * W + 1024 + cols)
    assert np.array_equal(ref.out_color[rs, 3], a.out_color[rs, 3])
    assert np.array_equal(ref.diagnostics["ray_count"][rs], a.diagnostics["ray_count"][rs])
    n = np.maximum(ref.out_color[rs, 3:4], 1)
    assert np.abs(ref.out_color[rs, :3] / n - a.out_color[rs, :3] / n).max() <= RGB_TOL


@pytest.mark.parametrize("name,count,spp", [("final", 0, 8), ("stress", 10000, 4)])
def test_full_frame_decisions_match_the_oracle(rtb, oracle, ctx, name, count, spp):
    """Every pixel of a full 1920x1080 frame (config 3's camera, BVH + defocus, depth 50) at a few spp: per-pixel
    successful-sample counts and ray counts equal the oracle's exactly, RGB within tolerance — 2 M pixels of evidence
    that no path takes a different decision anywhere in the frame."""
    W, H = 1920, 1080
    scene = rtb.host.make_scene(name, max_bvh_depth=16, target_count=count)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    assert_parity(ref, got, exact=False)


def test_random_small_batches_agree_across_kernels_and_with_the_oracle(rtb, oracle, ctx):
    """Fuzz: 40 random small batches (ragged sizes, 0..20 spp, depth 1..12, interlacing, row ranges, jitter on / off,
    adaptive sample ranges, previous accumulation) — the three kernels and the oracle take the same decisions."""
    rng = np.random.default_rng(2026)
    scenes = [rtb.host.make_scene("three_spheres", max_bvh_depth=int(d)) for d in (0, 2)] + \
             [rtb.host.make_scene("final", max_bvh_depth=int(d)) for d in (3, 16)] + [rtb.host.make_mesh_scene(max_bvh_depth=4)]
    for case in range(40):
        scene = scenes[int(rng.integers(len(scenes)))]
        W, H = int(rng.integers(1, 41)), int(rng.integers(1, 25))
        spp = int(rng.integers(0, 21))
        spp_max = spp + int(rng.integers(0, 6)) if rng.random() < 0.4 else None
        divider = int(rng.integers(1, 4))
        offset = int(rng.integers(0, divider))
        rb = int(rng.integers(0, H)) if rng.random() < 0.4 else 0
        re = int(rng.integers(rb + 1, H + 1)) if rb or rng.random() < 0.2 else 0
        p = rtb.host.make_params(scene, W, H, spp, int(rng.integers(1, 13)), seed=int(rng.integers(1, 1000)),
                                 aperture=float(rng.choice([0.0, 0.1, 0.4])), jitter=bool(rng.random() < 0.8),
                                 slice_offset=offset, slice_divider=divider, row_begin=rb, row_end=re, spp_max=spp_max)
        p.sample_count_weight_extrema[0], p.sample_count_weight_extrema[1] = 0.5, 2.5
        prev_n = rng.integers(0, 4, W * H).astype(np.float32)
        inputs = ((rng.random((W * H, 4)) * prev_n[:, None]).astype(np.float32), (rng.random(W * H) * 3 * prev_n).astype(np.float32),
                  (rng.random((W * H, 3)) * prev_n[:, None]).astype(np.float32), (rng.random((W * H, 3)) * prev_n[:, None]).astype(np.float32))
        inputs[0][:, 3] = prev_n
        ref = oracle.Buffers(W, H)
        ref.in_color[:], ref.in_weight[:], ref.in_normal[:], ref.in_albedo[:] = inputs
        oracle.sample_batch(scene, p, ref)
        for kernel, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
            got = render_gpu(rtb, ctx, scene, p, W, H, kernel, inputs=inputs)
            try:
                assert np.array_equal(ref.out_color[:, 3], got.out_color[:, 3])
                assert np.array_equal(ref.diagnostics["ray_count"], got.diagnostics["ray_count"])
                if exact:
                    assert np.array_equal(ref.out_color, got.out_color) and np.array_equal(ref.out_weight, got.out_weight)
                else:
                    assert np.abs(ref.out_color[:, :3] - got.out_color[:, :3]).max() <= 1e-4 * max(1.0, float(ref.out_color[:, 3].max()))
            except AssertionError:
                raise AssertionError(f"case {case}: {scene.name} {W}x{H} spp {spp}..{spp_max} div {divider}/{offset} rows {rb}:{re} kernel {kernel}")


def test_random_worlds_of_every_entity_kind(rtb, oracle, ctx):
    """Fuzz over WORLDS: 24 random scenes mixing plain spheres, rotated / moving spheres, Rects, Boxes and triangles at random
    poses (rays start inside boxes, graze rects edge-on, hit moving entities), random materials, leaf sizes and apertures —
    the per-pixel kernel is bit-identical to the oracle, the megakernel takes the same decisions."""
    for seed in range(24):
        scene = rtb.host.make_random_placed_scene(seed, count=12 + 3 * (seed % 5), max_bvh_depth=[0, 2, 8, 16][seed % 4])
        W, H, spp = 48, 27, 6
        p = rtb.host.make_params(scene, W, H, spp, 12 if seed % 3 else 50, seed=seed + 1, aperture=0.15 if seed % 2 else 0.0)
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, p, ref)
        for kernel, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
            got = render_gpu(rtb, ctx, scene, p, W, H, kernel)
            try:
                assert_parity(ref, got, exact=exact)
            except AssertionError as e:
                raise AssertionError(f"random world {seed}, kernel {kernel}: {e}")


@pytest.mark.parametrize("depth,moving,aperture", [(16, True, 0.0), (2, False, 0.15), (0, True, 0.0)])
def test_fog_cornell_world_matches_the_oracle(rtb, oracle, ctx, depth, moving, aperture):
    """MaterialType.ProbabilisticVolume (SampleBatchJob.cs:194-303, 450-524; Material.cs:48-65,163-168): balls of fog and
    smoke (one of them moving), a medium inside the glass ball, an inert Box medium — entry / exit bookkeeping over the sorted
    list of all hits, injected exit hits, backwards containment rays.  The collect-all kernel (RTB_OPT_KERNEL = 1) is bit-identical
    to the oracle; the megakernel's media flavour (two pruned walks instead of the list of every hit, media.cuh) takes the same
    decisions — equal sample counts and ray counts in every pixel, sums equal up to the accumulation order."""
    W, H, spp = 96, 54, 16
    scene = rtb.host.make_cornell_scene(max_bvh_depth=depth, moving=moving, fog=True)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=aperture)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    plain = oracle.Buffers(W, H)
    oracle.sample_batch(rtb.host.make_cornell_scene(max_bvh_depth=depth, moving=moving), p, plain)
    assert np.abs(ref.rgb() - plain.rgb()).max() > 0.5                       # the media are visible
    for kernel, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
        got = render_gpu(rtb, ctx, scene, p, W, H, kernel)
        assert_parity(ref, got, exact=exact)


def test_media_flavour_takes_the_collect_all_kernels_decisions(rtb, ctx):
    """The megakernel's media flavour against the collect-all validator kernel on larger frames than the oracle affords (GPU
    against GPU): the fog Cornell box, the same with the camera inside a ball of fog (every camera ray starts with the containment
    test and its backwards ray), and a linear list (one leaf holding every entity: the candidate order inside a leaf)."""
    W, H, spp = 320, 180, 32
    for depth, inside in ((16, False), (16, True), (0, False)):
        scene = rtb.host.make_cornell_scene(max_bvh_depth=depth, moving=True, fog=True)
        if inside:
            scene.spheres["center"][2] = (2.775, 2.775, -8.0)
            scene.spheres["radius"][2] = 1.5
            scene = rtb.host.build_world(scene.spheres, [], scene.materials, depth, scene.camera, scene.environment, scene.focus_distance,
                                         placed=scene.placed)
        p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.05)
        ref = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
        got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        assert_parity(ref, got, exact=False)


def test_random_worlds_with_media(rtb, oracle, ctx):
    """Fuzz over random worlds of every entity kind in which two or three of the eight materials are participating media —
    overlapping, nested, moving and non-convex media (Rects, triangles: no injected exit), media worn by several entities,
    multi-entity leaves (shallow trees) and linear lists: the collect-all kernel is bit-identical to the oracle, the
    megakernel's media flavour takes the same decisions; on larger frames the two kernels are compared with each other."""
    for seed in range(16):
        scene = rtb.host.make_random_placed_scene(seed, count=10 + 3 * (seed % 5), max_bvh_depth=[0, 2, 8, 16][seed % 4], media=2 + seed % 2)
        W, H, spp = 48, 27, 6
        p = rtb.host.make_params(scene, W, H, spp, 12 if seed % 3 else 50, seed=seed + 1, aperture=0.15 if seed % 2 else 0.0)
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, p, ref)
        for kernel, exact in ((rtb.abi.KERNEL_SIMPLE, True), (rtb.abi.KERNEL_MEGA, False)):
            got = render_gpu(rtb, ctx, scene, p, W, H, kernel)
            try:
                assert_parity(ref, got, exact=exact)
            except AssertionError as e:
                raise AssertionError(f"random media world {seed}, kernel {kernel}: {e}")
    for seed in (3, 6, 9):
        scene = rtb.host.make_random_placed_scene(seed, count=20, max_bvh_depth=[2, 16, 0][seed % 3], media=3)
        W, H, spp = 256, 144, 24
        p = rtb.host.make_params(scene, W, H, spp, 50, seed=seed, aperture=0.1)
        ref = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE)
        got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        try:
            assert_parity(ref, got, exact=False)
        except AssertionError as e:
            raise AssertionError(f"random media world {seed} at {W}x{H}x{spp}: {e}")


def test_fog_world_white_noise_stream_and_camera_inside_a_medium(rtb, oracle, ctx):
    """The reference's xorshift32 stream through the media (ProbabilisticHit's draw and the isotropic direction are taken in
    stream order), with the camera INSIDE a ball of fog so that every camera ray starts with the containment test."""
    W, H, spp = 64, 36, 8
    scene = rtb.host.make_cornell_scene(max_bvh_depth=16, moving=True, fog=True)
    scene.spheres["center"][2] = (2.775, 2.775, -8.0)                        # the white fog ball now surrounds the camera
    scene.spheres["radius"][2] = 1.5
    scene = rtb.host.build_world(scene.spheres, [], scene.materials, 16, scene.camera, scene.environment, scene.focus_distance,
                                 placed=scene.placed)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    for noise, opt in ((oracle.NOISE_PHILOX, rtb.abi.NOISE_PHILOX), (oracle.NOISE_XORSHIFT, rtb.abi.NOISE_WHITE)):
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, p, ref, noise=noise)
        assert ref.diagnostics["ray_count"].min() >= spp
        ctx.upload(scene)
        ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_SIMPLE)
        ctx.set_option(rtb.abi.OPT_NOISE, opt)
        try:
            got = rtb.plugin.HostBuffers(W, H)
            ctx.sample_batch(p, got)
        finally:
            ctx.set_option(rtb.abi.OPT_NOISE, 0)
        assert_parity(ref, got, exact=True)


def test_black_fog_transmittance_on_the_gpu(rtb, ctx):
    """Beer-Lambert through a non-scattering medium in front of a white sky, straight from the kernel (no oracle): the centre
    pixel is exp(-density * chord) from outside and from inside the medium."""
    import test_oracle_kat as kat
    spp = 4096
    for camera_z, length in ((-6.0, 2.0), (-0.5, 1.5)):
        s = kat._fog_world(rtb, camera_z, density=0.8)
        p = rtb.host.make_params(s, 3, 3, spp, 50, jitter=False)
        b = render_gpu(rtb, ctx, s, p, 3, 3, rtb.abi.KERNEL_MEGA)
        want = np.exp(-0.8 * length)
        assert b.out_color[4, 3] == spp
        assert np.allclose(b.rgb()[1, 1], want, atol=4 * np.sqrt(want * (1 - want) / spp)), (camera_z, b.rgb()[1, 1], want)


def test_an_unused_volume_material_does_not_change_the_kernel(rtb, ctx):
    """Only media that some entity wears send a world to the collect-all kernel: a ProbabilisticVolume left unused in the
    material buffer keeps the megakernel (same bits as without it)."""
    W, H, spp = 64, 36, 8
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    a = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    extra = rtb.host.make_scene("final", max_bvh_depth=16)
    extra.materials = np.concatenate([extra.materials, np.array([rtb.host._material(rtb.abi.MATERIAL_PROBABILISTIC_VOLUME, (1, 1, 1), ior=2.0)],
                                                                dtype=rtb.abi.MATERIAL_DTYPE)])
    b = render_gpu(rtb, ctx, extra, p, W, H, rtb.abi.KERNEL_MEGA)
    assert a.out_color.tobytes() == b.out_color.tobytes() and a.out_weight.tobytes() == b.out_weight.tobytes()


def test_accumulator_range_is_loud(rtb, oracle, ctx):
    """rtb.h "Accumulation range": the megakernel's per-pixel sums are 64-bit fixed point with 32 fraction bits.  Emitters far
    brighter than any display range still add up like the reference's floats (2^20 <= sample < 2^25 takes the one-at-a-time
    path); a pixel whose batch total reaches 2^30 is written as NaN — never a wrapped, sign-flipped value."""
    W, H, spp = 32, 18, 64
    for emission, overflow in ((3.0e6, False), (2.5e7, True)):
        scene = rtb.host.make_scene("three_spheres", max_bvh_depth=2)
        scene.materials = scene.materials.copy()
        scene.materials["type"][:] = rtb.abi.MATERIAL_STANDARD     # black diffuse emitters: every sample that hits one is `emission`
        scene.materials["glossiness"][:] = 0.0
        scene.materials["metallic"][:] = 0.0
        scene.materials["albedo"][:] = 0.0
        scene.materials["emission"][:] = emission
        p = rtb.host.make_params(scene, W, H, spp, 8)
        ref = oracle.Buffers(W, H)
        oracle.sample_batch(scene, p, ref)
        got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
        assert np.array_equal(ref.out_color[:, 3], got.out_color[:, 3])
        big = ref.out_color[:, 0] >= 2.0 ** 30
        assert big.any() == overflow
        assert np.isnan(got.out_color[big, :3]).all()
        ok = ~big
        assert np.isfinite(got.out_color[ok]).all()
        np.testing.assert_allclose(got.out_color[ok, :3], ref.out_color[ok, :3], rtol=2e-6)
        assert (got.out_color[ok, :3] >= 0).all()


def test_media_hit_list_overflow_is_an_error(rtb, ctx):
    """The reference's HybridList<HitRecord> starts at 32 records and grows (HybridCollections.cs:65-71); the volume kernel keeps
    48.  A ray through more media boundaries than that must fail the batch (RTB_ERR_UNSUPPORTED), not render something else."""
    abi = rtb.abi
    n = 40                                   # 40 nested fog shells along every camera ray: 3 records each (entry, exit, injected exit)
    materials = np.array([rtb.host._material(abi.MATERIAL_PROBABILISTIC_VOLUME, (0.9, 0.9, 0.9), ior=0.01)], dtype=abi.MATERIAL_DTYPE)
    spheres = np.zeros(n, dtype=abi.SPHERE_DTYPE)
    for i in range(n):
        spheres[i] = ((0.0, 0.0, 0.0), 1.0 + 0.05 * i, 0, (0, 0, 0))
    cam = abi.Camera()
    cam.position[:] = (0.0, 0.0, -8.0)
    cam.target[:] = (0.0, 0.0, 0.0)
    cam.aperture = 0.0
    cam.vertical_fov = 20.0
    env = abi.Environment()
    env.sky_type = abi.SKY_GRADIENT
    env.sky_bottom_color[:] = (1, 1, 1)
    env.sky_top_color[:] = (0.5, 0.7, 1.0)
    scene = rtb.host.build_world(spheres, [], materials, 4, cam, env, 8.0, name="shells")
    p = rtb.host.make_params(scene, 16, 9, 2, 8)
    c = rtb.plugin.Context(0)
    try:
        c.upload(scene)
        with pytest.raises(rtb.plugin.RtbError) as e:
            c.sample_batch(p, rtb.plugin.HostBuffers(16, 9))
        assert e.value.code == abi.RTB_ERR_UNSUPPORTED
        # fewer shells fit: the same world with 10 of them renders
        scene = rtb.host.build_world(spheres[:10], [], materials, 4, cam, env, 8.0, name="shells")
        c.upload(scene)
        c.sample_batch(p, rtb.plugin.HostBuffers(16, 9))
    finally:
        c.close()


def _multi_devices():
    import torch

    n = torch.cuda.device_count()
    return [0, 1] if n >= 2 else [0, 0]          # one GPU: two contexts share it, the tiling logic is the same


@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned-in-place"])
def test_multi_device_host_batch_is_bit_identical(rtb, ctx, pinned):
    """rtb_multi_sample_batch: ONE call renders the frame with every device of the handle (row tiles balanced inside the plugin,
    no gather).  Every output equals the single-device render bit for bit (order-independent sums, Philox keyed by the global
    pixel index), for pageable host arrays (staged per device) and pinned ones (written in place by every device)."""
    W, H, spp = 96, 54, 16
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    one = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    m = rtb.plugin.MultiContext(_multi_devices())
    try:
        m.upload(scene)
        b = rtb.plugin.HostBuffers(W, H)
        if pinned:
            m.register_host_buffers(b)
        for balance in (1, 0):
            m.set_option(rtb.abi.OPT_BALANCE_TILES, balance)
            for _ in range(3):                           # the third batch runs on tiles corrected by measured kernel times
                for a in b.arrays():
                    a[...] = 0
                m.sample_batch(p, b)
                bounds, ms = m.tiles()
                assert bounds[0] == 0 and bounds[-1] == H and all(bounds[g + 1] > bounds[g] for g in range(2))
                assert all(t > 0 for t in ms)
                for x, y in ((one.out_color, b.out_color), (one.out_normal, b.out_normal), (one.out_albedo, b.out_albedo),
                             (one.out_weight, b.out_weight), (one.diagnostics["ray_count"], b.diagnostics["ray_count"])):
                    assert x.tobytes() == y.tobytes()
        # interlaced rows and a row range compose with the tiling
        p2 = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1, slice_offset=1, slice_divider=3)
        p2.row_begin, p2.row_end = 5, 40
        ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_MEGA)
        ref = rtb.plugin.HostBuffers(W, H)
        ctx.sample_batch(p2, ref)
        for a in b.arrays():
            a[...] = 0
        m.sample_batch(p2, b)
        assert ref.out_color.tobytes() == b.out_color.tobytes()
        # the token reaches every device
        cancel = np.ones(1, np.uint8)
        with pytest.raises(rtb.plugin.RtbError) as e:
            m.sample_batch(p, b, cancel=cancel)
        assert e.value.code == rtb.abi.RTB_ERR_CANCELLED
        if pinned:
            m.unregister_host_buffers(b)
    finally:
        m.close()


def test_multi_device_batch_on_device_buffers(rtb, ctx):
    """rtb_multi_sample_batch_device: the accumulation buffers live on ONE device; the others write their row tiles into them
    over peer access (NVLink), ordered on the owner's stream — no gather, no collective.  Bit-identical to one device."""
    import torch

    W, H, spp = 96, 54, 16
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    one = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    m = rtb.plugin.MultiContext(_multi_devices())
    try:
        m.upload(scene)
        dev = torch.device("cuda:0")
        n = W * H
        z = lambda c: torch.zeros(n, c, device=dev, dtype=torch.float32)   # noqa: E731
        inp = [z(4), torch.zeros(n, device=dev), z(3), z(3)]
        out = [z(4), torch.zeros(n, device=dev), z(3), z(3)]
        diag = z(4)
        bufs = rtb.plugin.device_buffers_struct(*inp, *out, diag)
        stream = torch.cuda.current_stream(dev)
        for _ in range(3):
            for t in out:
                t.zero_()
            m.sample_batch_device(p, bufs, owner_index=0, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            assert out[0].cpu().numpy().tobytes() == one.out_color.tobytes()
            assert out[2].cpu().numpy().tobytes() == one.out_normal.tobytes()
            assert out[3].cpu().numpy().tobytes() == one.out_albedo.tobytes()
            assert out[1].cpu().numpy().tobytes() == one.out_weight.tobytes()
            assert np.array_equal(diag.cpu().numpy()[:, 0], one.diagnostics["ray_count"])
        bounds, ms = m.tiles()
        assert bounds[0] == 0 and bounds[-1] == H and all(t > 0 for t in ms)
    finally:
        m.close()


def test_peer_process_frame_mapping(rtb, ctx):
    """rtb_device_alloc / rtb_ipc_export: memory another rank process maps to write its tile into (the export side; the
    open side needs a second process and is exercised by bench.py --gpus N)."""
    ptr = ctx.device_alloc(1 << 20)
    try:
        h = ctx.ipc_export(ptr)
        assert len(h) == 64 and any(h)
    finally:
        ctx.device_free(ptr)


@pytest.mark.parametrize("name,depth", [("three_spheres", 0), ("final", 16)])
def test_white_furnace_on_the_gpu(rtb, ctx, name, depth):
    """tests/test_oracle_kat.py::test_white_furnace_radiance_is_exactly_one, on every kernel: albedo 1 under a sky of radiance 1
    gives pixel sums EQUAL to the sample counts — a check that needs no oracle at all."""
    from test_oracle_kat import furnace_scene

    W, H, spp = 96, 54, 16
    scene = furnace_scene(rtb, name, depth)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    for kernel in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
        b = render_gpu(rtb, ctx, scene, p, W, H, kernel)
        n = b.out_color[:, 3]
        assert n.max() == spp and n.min() >= spp - 2
        for c in range(3):
            assert np.array_equal(b.out_color[:, c], n)
    ctx.set_option(rtb.abi.OPT_KERNEL, rtb.abi.KERNEL_MEGA)
    ctx.set_option(rtb.abi.OPT_MATH, rtb.abi.MATH_FAST)      # the fast-arithmetic build conserves energy exactly too
    try:
        b = rtb.plugin.HostBuffers(W, H)
        ctx.sample_batch(p, b)
        n = b.out_color[:, 3]
        for c in range(3):
            assert np.array_equal(b.out_color[:, c], n)
    finally:
        ctx.set_option(rtb.abi.OPT_MATH, rtb.abi.MATH_PARITY)


def test_lambertian_albedo_series_on_the_gpu(rtb, ctx):
    """colour == a^(RayCount - 1) at one sample per pixel (see test_oracle_kat.py::test_lambertian_albedo_series)."""
    from test_oracle_kat import furnace_scene

    W, H = 128, 72
    scene = furnace_scene(rtb, "final", 16, albedo=0.5, lambertian_only=True)
    p = rtb.host.make_params(scene, W, H, 1, 30, aperture=0.0)
    for kernel in (rtb.abi.KERNEL_SIMPLE, rtb.abi.KERNEL_MEGA):
        b = render_gpu(rtb, ctx, scene, p, W, H, kernel)
        ok = b.out_color[:, 3] == 1
        assert ok.mean() > 0.99
        rays = b.diagnostics["ray_count"][ok].astype(np.int32)
        want = np.ldexp(np.float32(1.0), -(rays - 1)).astype(np.float32)
        for c in range(3):
            assert np.array_equal(b.out_color[ok, c], want)


def test_fast_math_build_agrees_statistically(rtb, ctx):
    """RTB_OPT_MATH = 1 (fast_kernels.cu): same estimator, same random numbers, approximate arithmetic.  Against the parity
    build at 64 spp: most pixels identical to 1e-4, the image mean within 1e-4, sample counts equal almost everywhere."""
    W, H, spp = 160, 90, 64
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    strict = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    ctx.set_option(rtb.abi.OPT_MATH, rtb.abi.MATH_FAST)
    try:
        fast = rtb.plugin.HostBuffers(W, H)
        ctx.sample_batch(p, fast)
    finally:
        ctx.set_option(rtb.abi.OPT_MATH, rtb.abi.MATH_PARITY)
    assert fast.out_color.tobytes() != strict.out_color.tobytes()           # it really is another build
    d = np.abs(strict.rgb() - fast.rgb()).reshape(-1, 3).max(axis=1)
    assert (d > 1e-4).mean() < 0.10
    assert abs(float(strict.rgb().mean()) - float(fast.rgb().mean())) < 1e-4
    assert (strict.out_color[:, 3] != fast.out_color[:, 3]).mean() < 0.01


# ---- RTB_OPT_RETREE: another topology over the host tree's leaves (csrc/retree.hpp; CPU side: tests/test_retree.py) ----------
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["simple", "mega"])
@pytest.mark.parametrize("case", [("final", 16, 0, 256, 144, 32, 0.1), ("final", 3, 0, 96, 54, 16, 0.1), ("stress", 16, 3000, 160, 90, 16, 0.1),
                                  ("final", 16, 0, 1920, 1080, 8, 0.1), ("stress", 16, 10000, 1024, 1024, 4, 0.1)],
                         ids=lambda c: f"{c[0]}-bvh{c[1]}-{c[3]}x{c[4]}x{c[5]}")
def test_retree_does_not_change_a_bit_on_sphere_worlds(rtb, ctx, case, kernel):
    """Every output word of a batch — colour sums, AOVs, sample counts, ray counts — is the same whether the device walks the
    host's topology (RTB_OPT_RETREE = 0) or the re-built one (1, the default): the reference's candidates are the leaves whose
    own box is hit, whatever lies above them."""
    name, depth, target, W, H, spp, ap = case
    if kernel == "simple" and W * H > 100000:
        pytest.skip("full-size frames through the megakernel only")
    scene = rtb.host.make_scene(name, max_bvh_depth=depth, target_count=target)
    assert rtb.plugin.retree_bvh(scene.nodes) is not None
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=ap)
    k = {"simple": rtb.abi.KERNEL_SIMPLE, "mega": rtb.abi.KERNEL_MEGA}[kernel]
    out = []
    try:
        for mode in (0, 1):
            ctx.set_option(rtb.abi.OPT_RETREE, mode)
            out.append(render_gpu(rtb, ctx, scene, p, W, H, k))
    finally:
        ctx.set_option(rtb.abi.OPT_RETREE, 1)
    a, b = out
    assert np.array_equal(a.out_color, b.out_color) and np.array_equal(a.out_normal, b.out_normal)
    assert np.array_equal(a.out_albedo, b.out_albedo) and np.array_equal(a.out_weight, b.out_weight)
    assert np.array_equal(a.diagnostics["ray_count"], b.diagnostics["ray_count"])
    assert a.out_color[:, 3].sum() > 0


@pytest.mark.gpu
def test_retree_walks_fewer_boxes(rtb, ctx):
    """The point of it: the instrumented build counts the walk that ran — fewer box tests per ray through the re-built tree."""
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, 160, 90, 16, 50, aperture=0.1)
    per_ray = []
    ctx.set_option(rtb.abi.OPT_COUNTERS, 1)
    try:
        for mode in (0, 1):
            ctx.set_option(rtb.abi.OPT_RETREE, mode)
            render_gpu(rtb, ctx, scene, p, 160, 90, rtb.abi.KERNEL_MEGA)
            cnt = ctx.counters()
            per_ray.append(cnt["node_tests"] / cnt["rays"])
    finally:
        ctx.set_option(rtb.abi.OPT_RETREE, 1)
        ctx.set_option(rtb.abi.OPT_COUNTERS, 0)
    assert per_ray[1] < 0.95 * per_ray[0], per_ray


@pytest.mark.gpu
@pytest.mark.parametrize("world", ["mesh", "cornell"])
def test_retree_opt_in_for_the_wider_worlds_keeps_oracle_parity(rtb, oracle, ctx, world):
    """RTB_OPT_RETREE = 2 (triangles, placed entities): the decisions and the image still match the oracle on the test sizes;
    the default (1) leaves these worlds on the host's topology."""
    scene = rtb.host.make_mesh_scene(max_bvh_depth=16, subdivisions=2) if world == "mesh" else rtb.host.make_cornell_scene(max_bvh_depth=16)
    W, H, spp = 96, 54, 8
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.0)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    try:
        ctx.set_option(rtb.abi.OPT_RETREE, 2)
        got = render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA)
    finally:
        ctx.set_option(rtb.abi.OPT_RETREE, 1)
    assert_parity(ref, got, exact=False)


def _coincident_triangles_scene(rtb, max_bvh_depth):
    """Every hit on the tilted quad is an EXACT TIE: its two triangles exist twice, at the same vertices, with different albedos.
    The reference's winner is decided by its candidate order (which leaf is visited later, SampleBatchJob.cs:420-475); the
    image shows which copy won."""
    abi, host = rtb.abi, rtb.host
    materials = np.zeros(4, dtype=abi.MATERIAL_DTYPE)
    for i, albedo in enumerate([(0.9, 0.1, 0.1), (0.1, 0.9, 0.1), (0.1, 0.1, 0.9), (0.8, 0.8, 0.8)]):
        materials[i]["type"], materials[i]["albedo"] = abi.MATERIAL_STANDARD, albedo
    q = [(-3.0, 0.0, -3.0), (3.0, 0.6, -3.0), (3.0, 1.1, 3.0), (-3.0, 0.4, 3.0)]
    tris = [host.make_triangle(q[0], q[2], q[1], 0), host.make_triangle(q[0], q[3], q[2], 0),
            host.make_triangle(q[0], q[2], q[1], 1), host.make_triangle(q[0], q[3], q[2], 1),
            host.make_triangle(q[0], q[2], q[1], 2),                                           # a third copy of one of them
            host.make_triangle((-1.0, 2.0, -1.0), (1.0, 2.5, 0.0), (0.0, 2.2, 1.0), 3)]        # something else in the tree
    spheres = np.zeros(1, dtype=abi.SPHERE_DTYPE)
    spheres[0] = ((0.5, 1.6, 0.3), 0.5, 3, (0, 0, 0))
    cam = abi.Camera()
    cam.position[:] = (0.5, 7.0, -6.0)
    cam.target[:] = (0.0, 0.5, 0.0)
    cam.aperture = 0.0
    cam.vertical_fov = 50.0
    env = abi.Environment()
    env.sky_type = abi.SKY_GRADIENT
    env.sky_bottom_color[:] = (1.0, 1.0, 1.0)
    env.sky_top_color[:] = (0.5, 0.7, 1.0)
    return host.build_world(spheres, np.array(tris, dtype=abi.TRIANGLE_DTYPE), materials, max_bvh_depth, cam, env, 9.0, name="ties")


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [16, 1, 0])
def test_exact_distance_ties_go_to_the_references_candidate_order(rtb, oracle, ctx, depth):
    """Coincident triangles of different colours (every hit a tie): the triangle flavour picks the oracle's winner — separate
    leaves (depth 16), shared leaves (depth 1), one leaf (a linear list) — and the image does not depend on the walk's topology."""
    scene = _coincident_triangles_scene(rtb, depth)
    W, H, spp = 96, 54, 8
    p = rtb.host.make_params(scene, W, H, spp, 12, aperture=0.0)
    ref = oracle.Buffers(W, H)
    oracle.sample_batch(scene, p, ref)
    assert ref.out_albedo.std(axis=0).max() > 0.05          # the quad is in view
    out = []
    try:
        for mode in (0, 1):
            ctx.set_option(rtb.abi.OPT_RETREE, mode)
            out.append(render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA))
            assert_parity(ref, out[-1], exact=False)
            assert_parity(ref, render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_SIMPLE), exact=True)
    finally:
        ctx.set_option(rtb.abi.OPT_RETREE, 1)
    assert np.array_equal(out[0].out_color, out[1].out_color) and np.array_equal(out[0].out_albedo, out[1].out_albedo)


@pytest.mark.gpu
def test_retree_does_not_change_a_bit_on_the_mesh_world(rtb, ctx):
    """Spheres + triangles: RTB_OPT_RETREE 0 and 1 (the default applies to this world) give the same words."""
    scene = rtb.host.make_mesh_scene(max_bvh_depth=16, subdivisions=3)
    assert rtb.plugin.retree_bvh(scene.nodes) is not None
    W, H, spp = 640, 360, 16
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.0)
    out = []
    try:
        for mode in (0, 1):
            ctx.set_option(rtb.abi.OPT_RETREE, mode)
            out.append(render_gpu(rtb, ctx, scene, p, W, H, rtb.abi.KERNEL_MEGA))
    finally:
        ctx.set_option(rtb.abi.OPT_RETREE, 1)
    a, b = out
    assert np.array_equal(a.out_color, b.out_color) and np.array_equal(a.out_normal, b.out_normal)
    assert np.array_equal(a.out_albedo, b.out_albedo) and np.array_equal(a.diagnostics["ray_count"], b.diagnostics["ray_count"])


# ---- mid-flight cancellation, last in the suite: these are the only tests that stop a running 100 ms kernel, and the one
# place where a run ever ended with a sticky CUDA error — everything else has run by then ----------------------
def test_cancellation_token_set_mid_flight(rtb):
    """CancellationToken (SampleBatchJob.cs:61; Raytracer.cs:189-192 flips it through a raw pointer while the job runs): the
    batch is ONE launch whether or not a token is passed; the kernel polls the context's mapped flag when a warp claims a tile.
    A token set from another thread in the middle of a ~120 ms batch ends the call within 5 ms with RTB_ERR_CANCELLED; a live
    token that is never set costs nothing measurable (same kernel, same launch)."""
    import threading
    import time

    abi = rtb.abi
    W, H, spp = 1920, 1080, 256
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    p = rtb.host.make_params(scene, W, H, spp, 50, aperture=0.1)
    c = rtb.plugin.Context(0)
    try:
        c.upload(scene)
        b = rtb.plugin.HostBuffers(W, H, diagnostics=False)
        c.register_host_buffers(b)
        c.sample_batch(p, b)                                   # warm-up, and the time of an undisturbed batch
        t_plain = []
        for _ in range(2):
            t = time.perf_counter()
            c.sample_batch(p, b)
            t_plain.append(time.perf_counter() - t)
        cancel = np.zeros(1, np.uint8)
        t_token = []
        for _ in range(2):
            t = time.perf_counter()
            c.sample_batch(p, b, cancel=cancel)                # live token, never set
            t_token.append(time.perf_counter() - t)
        assert min(t_token) <= min(t_plain) * 1.02, (t_token, t_plain)
        full = min(t_plain)
        for delay in (0.010, 0.040):
            cancel[0] = 0
            set_at = [0.0]

            def fire():
                time.sleep(delay)
                set_at[0] = time.perf_counter()
                cancel[0] = 1

            th = threading.Thread(target=fire)
            th.start()
            with pytest.raises(rtb.plugin.RtbError) as e:
                c.sample_batch(p, b, cancel=cancel)
            returned = time.perf_counter()
            th.join()
            assert e.value.code == abi.RTB_ERR_CANCELLED
            assert returned - set_at[0] < 0.005, f"cancel took {1e3 * (returned - set_at[0]):.2f} ms"
            assert returned - set_at[0] + delay < 0.8 * full          # it really stopped early
        cancel[0] = 0
        c.sample_batch(p, b, cancel=cancel)                    # and the context is usable afterwards
        assert b.out_color[:, 3].min() > 0
    finally:
        c.close()


def test_cancellation_stops_the_kernels_whose_warps_run_in_step(rtb):
    """The placed-entity and media flavours synchronise their CTA's warps twice per trip and leave the loop by consensus
    (sample_kernels.cuh: kPhased): a token set mid-flight must still end the batch at once — no warp may wait at a barrier
    for one that has left —, and the context must render the full frame afterwards.  The media world also reads a status word
    back after every batch, and the placed world is rendered through PAGEABLE host arrays here (staged copies): neither
    read-back may keep the calling thread from watching the token while the kernel runs."""
    import threading
    import time

    abi = rtb.abi
    W, H, spp = 1920, 1080, 64
    for fog in (False, True):
        scene = rtb.host.make_cornell_scene(max_bvh_depth=16, fog=fog)
        p = rtb.host.make_params(scene, W, H, spp, 50)
        c = rtb.plugin.Context(0)
        try:
            c.upload(scene)
            b = rtb.plugin.HostBuffers(W, H, diagnostics=False)
            if fog:
                c.register_host_buffers(b)
            c.sample_batch(p, b)
            t = time.perf_counter()
            c.sample_batch(p, b)
            full = time.perf_counter() - t
            want = b.out_color.copy()
            cancel = np.zeros(1, np.uint8)
            set_at = [0.0]

            def fire():
                time.sleep(0.25 * full)
                set_at[0] = time.perf_counter()
                cancel[0] = 1

            th = threading.Thread(target=fire)
            th.start()
            with pytest.raises(rtb.plugin.RtbError) as e:
                c.sample_batch(p, b, cancel=cancel)
            returned = time.perf_counter()
            th.join()
            assert e.value.code == abi.RTB_ERR_CANCELLED
            assert returned - set_at[0] < 0.02, f"cancel took {1e3 * (returned - set_at[0]):.2f} ms"
            cancel[0] = 0
            c.sample_batch(p, b, cancel=cancel)
            assert np.array_equal(b.out_color, want)               # same bits as the undisturbed batch
        finally:
            c.close()
