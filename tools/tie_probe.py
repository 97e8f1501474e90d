"""Evidence for DESIGN 3.1a (why RTB_OPT_RETREE = 2 is opt-in): rays that land on an edge shared by two triangles of the mesh world are
accepted by BOTH triangles about one time in five, and a third of those pairs report EXACTLY the same distance — an exact tie, whose
winner the reference leaves to its (unstable) sort and a pruned walk to its visiting order.  CPU only (oracle triangle test).
Measured here: 20000 rays aimed at shared edges -> 12048 hit by one triangle, 4456 by both, 1494 of those with equal distance."""
import sys, ctypes as C
import numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib as O
rtb = O.rtb; abi = rtb.abi
scene = rtb.host.make_mesh_scene(max_bvh_depth=16, subdivisions=3)
tris = scene.triangles
L = O.lib(); f3 = abi.f32x3
v0 = tris["v0"].astype(np.float64); e1 = tris["edge1"].astype(np.float64); e2 = tris["edge2"].astype(np.float64)
verts = np.stack([v0, v0 + e1, v0 + e2], axis=1)   # [n, 3, 3]
# neighbours sharing an edge: brute force on rounded vertex keys
key = lambda p: tuple(np.round(p, 5))
edge_map = {}
for i in range(len(tris)):
    ks = [key(verts[i, j]) for j in range(3)]
    for a, b in ((0, 1), (1, 2), (2, 0)):
        edge_map.setdefault(frozenset((ks[a], ks[b])), []).append((i, a, b))
pairs = [v for v in edge_map.values() if len(v) == 2]
print("shared edges", len(pairs))
rng = np.random.default_rng(7)
both = 0; equal = 0; trials = 0; one = 0
for trial in range(20000):
    (i, a, b), (j, _, _) = pairs[rng.integers(len(pairs))]
    w = rng.uniform(0.05, 0.95)
    P = (verts[i, a] * (1 - w) + verts[i, b] * w)
    n = np.cross(e1[i], e2[i]); n /= np.linalg.norm(n)
    o = (P + n * rng.uniform(0.5, 5.0) + rng.normal(size=3) * 0.5).astype(np.float32)
    d = (P - o.astype(np.float64)); d = (d / np.linalg.norm(d)).astype(np.float32)
    res = []
    for t in (i, j):
        dist = C.c_float(0); pt = f3(); nm = f3()
        tri = tris[t:t + 1]
        res.append(dist.value if L.oracle_triangle_hit(tri.ctypes.data, f3(*o), f3(*d), C.byref(dist), pt, nm) and True else None)
        if res[-1] is not None: res[-1] = dist.value
    trials += 1
    if res[0] is not None and res[1] is not None:
        both += 1
        if res[0] == res[1]: equal += 1
    elif res[0] is not None or res[1] is not None:
        one += 1
print("rays aimed at a shared edge:", trials, "hit by exactly one:", one, "by both:", both, "of which with EQUAL distance:", equal)
