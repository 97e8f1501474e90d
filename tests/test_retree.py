"""csrc/retree.hpp through rtb_retree_bvh (host-side, no GPU): the tree the device walks under RTB_OPT_RETREE keeps the
reference's leaves, and — the claim the option rests on — the reference's candidate set (every box on the chain from the root
passes AxisAlignedCuboid.Hit, SampleBatchJob.cs:403-447) is the set of leaves whose OWN box is hit, so it is the same through
any topology over those leaves.  Checked numerically with the oracle's slab test on random and adversarial rays."""
import numpy as np
import pytest


def _children(nodes):
    return [(int(n["left"]), int(n["right"])) if n["first_entity"] < 0 else None for n in nodes]


def _leaves(nodes):
    """{(first_entity, entity_count): (bounds_min, bounds_max)} of the non-empty leaves reachable from the root, and the depth."""
    out, deepest, stack = {}, 0, [(0, 0)]
    while stack:
        i, d = stack.pop()
        n = nodes[i]
        if n["first_entity"] >= 0:
            if n["entity_count"] > 0:
                out[(int(n["first_entity"]), int(n["entity_count"]))] = (n["bounds_min"].copy(), n["bounds_max"].copy())
                deepest = max(deepest, d)
        else:
            stack += [(int(n["left"]), d + 1), (int(n["right"]), d + 1)]
    return out, deepest


def _slab_hits(nodes, origins, directions):
    """AxisAlignedCuboid.Hit for every (node, ray) in float32 — the arithmetic of the oracle's AabbHit (checked against it in
    test_slab_restatement_matches_the_oracle): (bound - o) * (1 / d) with NaN reciprocals -> +inf, math.min / max drop NaNs."""
    o = origins.astype(np.float32)[None, :, :]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv = (np.float32(1.0) / directions.astype(np.float32))[None, :, :]
        inv = np.where(np.isnan(inv), np.float32(np.inf), inv)
        t0 = (nodes["bounds_min"].astype(np.float32)[:, None, :] - o) * inv
        t1 = (nodes["bounds_max"].astype(np.float32)[:, None, :] - o) * inv
        lo, hi = np.fmin(t0, t1), np.fmax(t0, t1)
        t_min = np.fmax(np.float32(0.0), np.fmax(np.fmax(lo[..., 0], lo[..., 1]), lo[..., 2]))
        t_max = np.fmin(np.fmin(hi[..., 0], hi[..., 1]), hi[..., 2])
        return t_min < t_max


def _chain_hits(nodes, own):
    """own[node, ray] -> per non-empty leaf key: every box from the root down to the leaf is hit."""
    out, stack = {}, [(0, own[0])]
    while stack:
        i, alive = stack.pop()
        n = nodes[i]
        if n["first_entity"] >= 0:
            if n["entity_count"] > 0:
                out[(int(n["first_entity"]), int(n["entity_count"]))] = alive
        else:
            for c in (int(n["left"]), int(n["right"])):
                stack.append((c, alive & own[c]))
    return out


def _rays(scene, nodes, rng, count):
    """Random rays through the world plus the adversarial ones: directions with exact zero components, origins exactly on
    leaf-box planes (0 * inf = NaN slabs), origins 0.001 off a sphere (bounce rays), rays grazing box edges."""
    lo, hi = nodes[0]["bounds_min"], nodes[0]["bounds_max"]
    centre, ext = (lo + hi) / 2, np.minimum((hi - lo) / 2, 30.0)
    o = centre + rng.uniform(-1.2, 1.2, (count, 3)) * ext
    d = rng.normal(size=(count, 3))
    leaves = [n for n in nodes if n["first_entity"] >= 0 and n["entity_count"] > 0]
    k = count // 4
    for i in range(k):                      # zero components, origin on a leaf's plane along that axis
        lf = leaves[rng.integers(len(leaves))]
        axis = rng.integers(3)
        d[i, axis] = 0.0 if i % 2 else -0.0
        if i % 3 == 0:
            d[i, (axis + 1) % 3] = 0.0
        o[i] = (lf["bounds_min"] + lf["bounds_max"]) / 2 + rng.uniform(-0.3, 0.3, 3)
        o[i, axis] = lf["bounds_min"][axis] if i % 4 < 2 else lf["bounds_max"][axis]
    for i in range(k, 2 * k):               # bounce rays: 0.001 off a sphere's surface, cosine-ish directions
        s = scene.spheres[rng.integers(len(scene.spheres))]
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        o[i] = np.asarray(s["center"], dtype=np.float64) + n * (abs(float(s["radius"])) + 0.001)
        d[i] = n + rng.normal(size=3) * 0.7
    for i in range(2 * k, 3 * k):           # aimed at a leaf-box corner: grazing its slabs
        lf = leaves[rng.integers(len(leaves))]
        corner = np.where(rng.integers(0, 2, 3) == 1, lf["bounds_max"], lf["bounds_min"])
        d[i] = corner - o[i]
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    return o.astype(np.float32), d.astype(np.float32)


def test_slab_restatement_matches_the_oracle(oracle, rtb):
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    rng = np.random.default_rng(5)
    o, d = _rays(scene, scene.nodes, rng, 64)
    own = _slab_hits(scene.nodes, o, d)
    L = oracle.lib()
    f3 = rtb.abi.f32x3
    for r in range(len(o)):
        for i in range(0, len(scene.nodes), 7):
            n = scene.nodes[i]
            assert bool(L.oracle_aabb_hit(f3(*n["bounds_min"]), f3(*n["bounds_max"]), f3(*o[r]), f3(*d[r]))) == bool(own[i, r])


@pytest.mark.parametrize("name,depth,target", [("final", 16, 0), ("final", 3, 0), ("final", 32, 0), ("stress", 16, 3000), ("three_spheres", 2, 0)])
def test_retree_keeps_the_leaves_and_bounds_what_lies_below(rtb, name, depth, target):
    scene = rtb.host.make_scene(name, max_bvh_depth=depth, target_count=target)
    new = rtb.plugin.retree_bvh(scene.nodes)
    assert new is not None
    ref_leaves, _ = _leaves(scene.nodes)
    new_leaves, deepest = _leaves(new)
    assert ref_leaves.keys() == new_leaves.keys()
    for k in ref_leaves:
        assert np.array_equal(ref_leaves[k][0], new_leaves[k][0]) and np.array_equal(ref_leaves[k][1], new_leaves[k][1])
    assert len(new) == 2 * len(new_leaves) - 1 and deepest <= 62
    for n in new:
        if n["first_entity"] < 0:
            l, r = new[n["left"]], new[n["right"]]
            assert np.array_equal(n["bounds_min"], np.minimum(l["bounds_min"], r["bounds_min"]))
            assert np.array_equal(n["bounds_max"], np.maximum(l["bounds_max"], r["bounds_max"]))
    again = rtb.plugin.retree_bvh(scene.nodes)
    assert new.tobytes() == again.tobytes()          # deterministic


@pytest.mark.parametrize("name,depth", [("final", 16), ("final", 4)])
def test_candidates_do_not_depend_on_the_topology(rtb, name, depth):
    """reference chain == the leaf's own box == the re-built tree's chain, for every leaf and ray."""
    scene = rtb.host.make_scene(name, max_bvh_depth=depth)
    new = rtb.plugin.retree_bvh(scene.nodes)
    rng = np.random.default_rng(11)
    o, d = _rays(scene, scene.nodes, rng, 1600)
    own_ref = _slab_hits(scene.nodes, o, d)
    own_new = _slab_hits(new, o, d)
    chain_ref = _chain_hits(scene.nodes, own_ref)
    chain_new = _chain_hits(new, own_new)
    leaf_index = {(int(n["first_entity"]), int(n["entity_count"])): i for i, n in enumerate(scene.nodes) if n["first_entity"] >= 0 and n["entity_count"] > 0}
    total = 0
    for k, i in leaf_index.items():
        assert np.array_equal(chain_ref[k], own_ref[i]), k
        assert np.array_equal(chain_new[k], own_ref[i]), k
        total += int(own_ref[i].sum())
    assert total > 1000          # the rays do hit boxes


def test_worlds_that_do_not_qualify_keep_the_hosts_topology(rtb):
    scene = rtb.host.make_scene("final", max_bvh_depth=16)
    nodes = scene.nodes.copy()
    leaf = next(i for i, n in enumerate(nodes) if n["first_entity"] >= 0 and n["entity_count"] > 0)
    nodes[leaf]["bounds_max"][1] = nodes[leaf]["bounds_min"][1]            # a flat leaf box: 0 * inf could tell the chains apart
    assert rtb.plugin.retree_bvh(nodes) is None
    nodes = scene.nodes.copy()
    inner = next(i for i, n in enumerate(nodes) if n["first_entity"] < 0 and i > 0)
    nodes[inner]["bounds_max"][0] = nodes[inner]["bounds_min"][0]          # an inner box that does not contain its children: the chain matters
    assert rtb.plugin.retree_bvh(nodes) is None
    nodes = scene.nodes.copy()
    nodes[0]["left"] = 0                                                    # not a tree
    assert rtb.plugin.retree_bvh(nodes) is None
    assert rtb.plugin.retree_bvh(rtb.host.make_scene("final", max_bvh_depth=0).nodes) is None    # a linear list is one leaf


def test_depth_is_bounded_for_a_degenerate_world(rtb):
    """Nested shells (every split peels one sphere off): the builder halves by count before the walk's stack could overflow."""
    n = 400
    spheres = np.zeros(n, dtype=rtb.abi.SPHERE_DTYPE)
    for i in range(n):
        spheres[i]["center"] = (0.0, 0.0, 0.0)
        spheres[i]["radius"] = 1.5 ** (i * 0.2) * (1 + 1e-3 * i)
    ordered, nodes = rtb.host.build_bvh(spheres, 16)
    new = rtb.plugin.retree_bvh(nodes)
    if new is not None:
        _, deepest = _leaves(new)
        assert deepest <= 62


@pytest.mark.parametrize("seed", range(6))
def test_candidates_do_not_depend_on_the_topology_random_worlds(rtb, seed):
    """The same claim on random sphere worlds: overlapping, nested, tiny and huge spheres, negative radii (hollow glass),
    random BVH depth limits (multi-entity leaves)."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(5, 300))
    spheres = np.zeros(n, dtype=rtb.abi.SPHERE_DTYPE)
    spheres["center"] = rng.normal(size=(n, 3)) * rng.choice([1.0, 10.0, 300.0])
    spheres["radius"] = np.exp(rng.uniform(-4, 3, n)) * rng.choice([1.0, 1.0, 1.0, -1.0], n)
    if seed % 2:
        spheres["radius"][0] = 1000.0
        spheres["center"][0] = (0.0, -1000.0, 0.0)
    ordered, nodes = rtb.host.build_bvh(spheres, int(rng.integers(1, 20)))
    new = rtb.plugin.retree_bvh(nodes)
    if new is None:           # fewer than two non-empty leaves
        assert sum(1 for x in nodes if x["first_entity"] >= 0 and x["entity_count"] > 0) < 2
        return

    class S:
        pass
    scene = S()
    scene.spheres = ordered
    o, d = _rays(scene, nodes, rng, 600)
    own_ref = _slab_hits(nodes, o, d)
    chain_ref = _chain_hits(nodes, own_ref)
    chain_new = _chain_hits(new, _slab_hits(new, o, d))
    for i, x in enumerate(nodes):
        if x["first_entity"] >= 0 and x["entity_count"] > 0:
            k = (int(x["first_entity"]), int(x["entity_count"]))
            assert np.array_equal(chain_ref[k], own_ref[i]) and np.array_equal(chain_new[k], own_ref[i]), k
    _, deepest = _leaves(new)
    assert deepest <= 62


def test_candidates_do_not_depend_on_the_topology_mesh_world(rtb):
    """The claim is about boxes, whatever the leaves hold: the mesh world (triangles + spheres, some leaves of several entities)."""
    scene = rtb.host.make_mesh_scene(max_bvh_depth=16, subdivisions=2)
    new = rtb.plugin.retree_bvh(scene.nodes)
    assert new is not None
    rng = np.random.default_rng(23)
    o, d = _rays(scene, scene.nodes, rng, 800)
    own_ref = _slab_hits(scene.nodes, o, d)
    chain_ref = _chain_hits(scene.nodes, own_ref)
    chain_new = _chain_hits(new, _slab_hits(new, o, d))
    checked = 0
    for i, x in enumerate(scene.nodes):
        if x["first_entity"] >= 0 and x["entity_count"] > 0:
            k = (int(x["first_entity"]), int(x["entity_count"]))
            assert np.array_equal(chain_ref[k], own_ref[i]) and np.array_equal(chain_new[k], own_ref[i]), k
            checked += 1
    assert checked > 100


def test_big_world_takes_the_binned_path(rtb):
    """More leaves than the optimiser takes (65 536): binned splits down to 256 leaves, no optimisation passes — same guarantees."""
    rng = np.random.default_rng(41)
    n = 70_000
    spheres = np.zeros(n, dtype=rtb.abi.SPHERE_DTYPE)
    spheres["center"] = rng.uniform(-200, 200, (n, 3)) * np.array([1.0, 0.05, 1.0])
    spheres["radius"] = rng.uniform(0.1, 0.6, n)
    ordered, nodes = rtb.host.build_bvh(spheres, 20)
    new = rtb.plugin.retree_bvh(nodes)
    assert new is not None
    ref_leaves, _ = _leaves(nodes)
    new_leaves, deepest = _leaves(new)
    assert ref_leaves.keys() == new_leaves.keys() and len(new) == 2 * len(new_leaves) - 1 and deepest <= 62
    inner = new[new["first_entity"] < 0]
    l, r = new[inner["left"]], new[inner["right"]]
    assert np.array_equal(inner["bounds_min"], np.minimum(l["bounds_min"], r["bounds_min"]))
    assert np.array_equal(inner["bounds_max"], np.maximum(l["bounds_max"], r["bounds_max"]))

    def area_sum(t):
        e = (t["bounds_max"] - t["bounds_min"]).astype(np.float64)[t["first_entity"] < 0]
        return float((e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]).sum())
    assert area_sum(new) < area_sum(nodes)          # fewer expected visits than the reference's median splits
    # big subtrees are built concurrently and spliced in the serial order: the same array for any thread count
    import os
    was = os.environ.get("RTB_BUILD_THREADS")
    try:
        for threads in ("1", "3"):
            os.environ["RTB_BUILD_THREADS"] = threads
            assert rtb.plugin.retree_bvh(nodes).tobytes() == new.tobytes()
    finally:
        if was is None:
            os.environ.pop("RTB_BUILD_THREADS", None)
        else:
            os.environ["RTB_BUILD_THREADS"] = was
