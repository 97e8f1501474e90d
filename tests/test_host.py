"""Host-side logic: interlacing series, row activation, sharding partition, job mirror."""
import numpy as np
import pytest


def test_space_filling_series(rtb):
    # Tools.SpaceFillingSeries (Tools.cs:101-124): a permutation of 0..n-1 starting at 0 that halves gaps
    assert rtb.host.space_filling_series(1) == [0]
    assert rtb.host.space_filling_series(2) == [0, 1]
    assert rtb.host.space_filling_series(4) == [0, 2, 1, 3]
    assert rtb.host.space_filling_series(8) == [0, 4, 2, 6, 1, 3, 5, 7]
    for n in (3, 8, 16, 32):
        s = rtb.host.space_filling_series(n)
        assert sorted(s) == list(range(n)) and s[0] == 0
    # literal restatement, quirk included: for some non-power-of-two lengths i * increment overshoots
    assert rtb.host.space_filling_series(5) == [0, 3, 2, 4, 6]


def test_view_basis_is_orthonormal_and_left_handed(rtb):
    v = rtb.host.make_view((0, 0, -2.25), (0, 0, 0), (0, 1, 0), 60.0, 16 / 9, 0.2, 3.0)
    f, u, r = (np.array(tuple(x)) for x in (v.forward, v.up, v.right))
    assert np.allclose([f @ u, f @ r, u @ r], 0, atol=1e-6)
    assert np.allclose([np.linalg.norm(f), np.linalg.norm(u), np.linalg.norm(r)], 1, atol=1e-6)
    assert np.allclose(f, (0, 0, -1)) and np.allclose(r, (1, 0, 0))    # Forward = normalize(origin - lookAt), View.cs:24
    assert abs(v.lens_radius - 0.1) < 1e-7
    assert np.allclose(np.linalg.norm(tuple(v.vertical)), 2 * np.tan(np.radians(30)) * 3.0, rtol=1e-5)


def test_make_params_mirrors_schedule_sample(rtb):
    scene = rtb.host.make_scene("three_spheres")
    p = rtb.host.make_params(scene, 400, 225, 4, 8)
    assert (p.size[0], p.size[1]) == (400.0, 225.0)
    assert (p.slice_offset, p.slice_divider, p.seed) == (0, 1, 1)
    assert tuple(p.sample_count_range) == (4, 4) and p.trace_depth == 8 and p.sub_pixel_jitter == 1
    assert p.environment.sky_type == rtb.abi.SKY_GRADIENT


def test_job_mirror_builds_the_same_params(rtb):
    from importlib import import_module

    job = import_module("raytracing-in-one-weekend_b200.job")
    scene = rtb.host.make_scene("three_spheres")
    view, _ = rtb.host.view_for(scene, 64, 36)
    j = job.SampleBatchJob(Size=(64, 36), SliceOffset=1, SliceDivider=2, Seed=7, View=view, Environment=scene.environment,
                           SampleCountRange=(2, 5), TraceDepth=9, SubPixelJitter=False, SampleCountWeightExtrema=(0.5, 2.0))
    p = j.params()
    assert (p.size[0], p.size[1], p.slice_offset, p.slice_divider, p.seed) == (64.0, 36.0, 1, 2, 7)
    assert tuple(p.sample_count_range) == (2, 5) and p.trace_depth == 9 and p.sub_pixel_jitter == 0
    assert tuple(p.sample_count_weight_extrema) == (0.5, 2.0)
    assert bytes(p.view) == bytes(view)


def test_row_tiles_partition(rtb):
    from importlib import import_module

    sh = import_module("raytracing-in-one-weekend_b200.sharding")
    for h, w in [(1080, 8), (1080, 7), (225, 2), (5, 4), (2160, 8)]:
        tiles = sh.row_tiles(h, w)
        assert tiles[0][0] == 0 and tiles[-1][1] == h
        assert all(tiles[i][1] == tiles[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in tiles) - min(e - b for b, e in tiles) <= 1
    cost = np.concatenate([np.full(100, 1.0), np.full(100, 9.0)])
    tiles = sh.balanced_row_tiles(cost, 4)
    assert tiles[0][0] == 0 and tiles[-1][1] == 200 and all(e > b for b, e in tiles)
    sums = [cost[b:e].sum() for b, e in tiles]
    assert max(sums) / min(sums) < 1.15


@pytest.mark.parametrize("name,depth", [("final", 16), ("final", 0), ("final", 3), ("three_spheres", 2), ("three_spheres", 0)])
def test_device_layout_collapses_small_subtrees(rtb, name, depth):
    """rtb_describe_scene (host-side half of rtb_upload_scene): the device tree is the host's BVH with
    subtrees of <= leaf_spheres spheres collapsed into one leaf; every sphere the BVH references
    lands in exactly one leaf slot for every setting."""
    scene = rtb.host.make_scene(name, max_bvh_depth=depth)
    nodes = scene.nodes
    n_ref_spheres = int(nodes["entity_count"][nodes["first_entity"] >= 0].sum())
    n_ref_inner = int((nodes["first_entity"] < 0).sum())
    one = rtb.plugin.describe_scene(scene, 1)
    assert one["device_spheres"] == n_ref_spheres and one["inner_nodes"] == n_ref_inner
    assert one["collapsed"] == 0 and one["chain_boxes"] == 0
    prev = one["inner_nodes"]
    for k in (2, 4, 8, 15):
        lay = rtb.plugin.describe_scene(scene, k)
        assert lay["device_spheres"] == n_ref_spheres
        assert lay["inner_nodes"] <= prev and lay["max_depth"] <= one["max_depth"]
        assert lay["leaves"] == lay["inner_nodes"] + 1
        prev = lay["inner_nodes"]
        if lay["collapsed"]:
            assert lay["chain_boxes"] > 0 and lay["max_leaf_spheres"] <= max(k, one["max_leaf_spheres"])
    if name == "final" and depth == 16:
        assert rtb.plugin.describe_scene(scene, 8)["inner_nodes"] < n_ref_inner // 3


def test_add_mesh_bakes_the_transform_and_rotates_normals(rtb):
    """AddMeshRuntimeEntitiesJob.cs:60-80: vertices -> transform(rigid, v * scale), normals -> rotate(rot, n), uvs pass through;
    the triangles equal the ones the Triangle ctors give for the transformed vertices."""
    host = rtb.host
    v = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)], np.float32)
    nrm = np.array([(0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1)], np.float32)
    uv = np.array([(0, 0), (1, 0), (0, 1), (0.5, 0.5)], np.float32)
    idx = [0, 1, 2, 0, 2, 3]
    q = host.quat_axis_angle((0, 1, 0), 90.0)
    tris, tuv = host.add_mesh(v, idx, 3, normals=nrm, uvs=uv, rotation=q, position=(10, 0, 0), scale=2.0)
    assert len(tris) == 2 and (tris["material"] == 3).all()
    # a quarter turn about +Y takes x to -z and z to x (rotate(q, v), right-handed algebra), then scale 2 and the offset
    np.testing.assert_allclose(tris[0]["v0"], (10, 0, 0), atol=1e-6)
    np.testing.assert_allclose(tris[0]["edge1"], (0, 0, -2), atol=1e-6)      # v2 - v1
    np.testing.assert_allclose(tris[0]["edge2"], (0, 2, 0), atol=1e-6)       # v3 - v1
    np.testing.assert_allclose(tris[0]["normals"][1], (0, 0, -1), atol=1e-6)  # (1, 0, 0) rotated
    np.testing.assert_allclose(tris[1]["edge2"], (2, 0, 0), atol=1e-6)       # (0, 0, 1) * 2 rotated
    assert np.array_equal(tuv[0], uv[[0, 1, 2]]) and np.array_equal(tuv[1], uv[[0, 2, 3]])
    # face-normal form: the ctor's normalize(cross(Data[1], Data[0]))
    flat, _ = host.add_mesh(v, idx, 0, rotation=q, position=(10, 0, 0), scale=2.0)
    want = host.make_triangle(tris[0]["v0"], tris[0]["v0"] + tris[0]["edge1"], tris[0]["v0"] + tris[0]["edge2"], 0)
    assert flat[0].tobytes() == want.tobytes()
    with pytest.raises(ValueError):
        host.add_mesh(v, [0, 1, 9], 0)                                        # index past the vertex array


def test_plugin_row_balancer_matches_the_python_partition(rtb):
    """rtb_balance_rows (the C++ balancer behind rtb_multi_sample_batch) against sharding.balanced_row_tiles, the partition the
    one-process-per-GPU driver uses: same bounds on random cost profiles, every tile non-empty, ranges respected."""
    from importlib import import_module

    sh = import_module("raytracing-in-one-weekend_b200.sharding")
    rng = np.random.default_rng(7)
    for height in (8, 27, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            if world > height:
                continue
            cost = rng.random(height) ** 3 + 1e-3
            cost[: height // 3] *= 0.05                     # cheap sky rows at one end, as in the final scene
            want = sh.balanced_row_tiles(cost, world)
            got = rtb.plugin.balance_rows(cost, 0, height, world)
            assert got == [b for b, _ in want] + [height]
            assert all(got[g + 1] > got[g] for g in range(world))
            sums = [cost[got[g]:got[g + 1]].sum() for g in range(world)]
            assert max(sums) <= cost.sum() / world + cost.max() + 1e-9
    # a sub-range, equal rows without a model, more devices than rows
    assert rtb.plugin.balance_rows(None, 10, 20, 2) == [10, 15, 20]
    assert rtb.plugin.balance_rows(np.ones(100), 40, 60, 4) == [40, 45, 50, 55, 60]
    b = rtb.plugin.balance_rows(None, 0, 2, 4)
    assert b[0] == 0 and b[-1] == 2 and all(b[g + 1] >= b[g] for g in range(4)) and sum(b[g + 1] - b[g] for g in range(4)) == 2


def test_threaded_bvh_build_is_identical_to_the_serial_one(rtb, monkeypatch):
    """rtbh_build_bvh_from_bounds builds large subtrees concurrently and splices them in the serial recursion's order: the
    node array and the BVH-ordered entity list do not depend on the thread count (BvhNodeData.cs:122-213 is one serial job)."""
    import ctypes as C

    rng = np.random.default_rng(3)
    n = 60000
    c = rng.normal(size=(n, 3)).astype(np.float32) * 5
    e = (np.abs(rng.normal(size=(n, 3))) * 0.05 + 1e-3).astype(np.float32)
    bounds = np.ascontiguousarray(np.hstack([c - e, c + e]), dtype=np.float32)

    def build():
        order = np.zeros(n, np.uint32)
        nodes = np.zeros(2 * n + 1, dtype=rtb.abi.BVH_NODE_DTYPE)
        count = C.c_size_t(0)
        assert rtb.host.lib().rtbh_build_bvh_from_bounds(bounds.ctypes.data, n, 32, order.ctypes.data, n, nodes.ctypes.data, len(nodes), C.byref(count)) == 0
        return order, nodes[: count.value].copy()

    monkeypatch.setenv("RTB_BUILD_THREADS", "1")
    o1, n1 = build()
    monkeypatch.setenv("RTB_BUILD_THREADS", "8")
    o8, n8 = build()
    assert np.array_equal(o1, o8) and n1.tobytes() == n8.tobytes()
    assert sorted(o1.tolist()) == list(range(n))
