"""Stand-in for the C# host (`Raytracer : MonoBehaviour`): scene construction, BVH build,
camera/View construction — thin ctypes wrappers over librtb_host.so (include/rtb_host.h).

The reference keeps these steps on the host (BASELINE.json north_star); they are the
producers of the sample job's inputs, not part of the hot path.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _abi as abi
from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.host_lib_path()
        if not os.path.exists(path):
            path = _build.build_host()
        L = C.CDLL(path)
        L.rtbh_random_init.argtypes = [C.POINTER(abi.URandom), C.c_uint32]
        L.rtbh_random_init.restype = None
        L.rtbh_random_next_state.argtypes = [C.POINTER(abi.URandom)]
        L.rtbh_random_next_state.restype = C.c_uint32
        L.rtbh_random_next_float.argtypes = [C.POINTER(abi.URandom)]
        L.rtbh_random_next_float.restype = C.c_float
        L.rtbh_scene_generate.argtypes = [
            C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(abi.SceneInfo)
        ]
        L.rtbh_scene_generate.restype = C.c_int
        L.rtbh_build_bvh.argtypes = [
            C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        ]
        L.rtbh_build_bvh.restype = C.c_int
        L.rtbh_build_bvh_from_bounds.argtypes = [
            C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
        ]
        L.rtbh_build_bvh_from_bounds.restype = C.c_int
        L.rtbh_sphere_bounds.argtypes = [C.c_void_p, C.c_void_p]
        L.rtbh_sphere_bounds.restype = None
        L.rtbh_triangle_bounds.argtypes = [C.c_void_p, C.c_void_p]
        L.rtbh_triangle_bounds.restype = None
        L.rtbh_placed_bounds.argtypes = [C.c_void_p, C.c_void_p]
        L.rtbh_placed_bounds.restype = None
        L.rtbh_make_triangle.argtypes = [abi.f32x3, abi.f32x3, abi.f32x3, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.rtbh_make_triangle.restype = None
        L.rtbh_add_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.rtbh_add_mesh.restype = C.c_int
        L.rtbh_make_view.argtypes = [
            abi.f32x3, abi.f32x3, abi.f32x3, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(abi.View)
        ]
        L.rtbh_make_view.restype = None
        L.rtbh_hit_world.argtypes = [
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, abi.f32x3, abi.f32x3, C.POINTER(C.c_float)
        ]
        L.rtbh_hit_world.restype = C.c_int
        L.rtbh_view_from_camera.argtypes = [
            C.POINTER(abi.Camera), C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_float,
            C.POINTER(abi.View), C.POINTER(C.c_float),
        ]
        L.rtbh_view_from_camera.restype = None
        L.rtbh_space_filling_series.argtypes = [C.c_int, C.POINTER(C.c_int32), C.c_size_t]
        L.rtbh_space_filling_series.restype = C.c_int
        _lib = L
    return _lib


class UnityRandom:
    """Unity.Mathematics.Random (xorshift32)."""

    def __init__(self, seed):
        self._r = abi.URandom()
        lib().rtbh_random_init(C.byref(self._r), seed)

    @property
    def state(self):
        return self._r.state

    def next_state(self):
        return lib().rtbh_random_next_state(C.byref(self._r))

    def next_float(self):
        return lib().rtbh_random_next_float(C.byref(self._r))


@dataclass
class Scene:
    """A flattened world ready for `Context.upload_scene`: BVH-ordered spheres, materials, nodes."""

    spheres: np.ndarray  # abi.SPHERE_DTYPE, BVH (leaf) order
    materials: np.ndarray  # abi.MATERIAL_DTYPE
    nodes: np.ndarray  # abi.BVH_NODE_DTYPE, root = 0
    camera: abi.Camera
    environment: abi.Environment
    info: abi.SceneInfo
    max_bvh_depth: int
    name: str = ""
    # mixed worlds (rtb_upload_world): the BVH-ordered entity list the leaves index, and the triangle array
    entities: np.ndarray = None  # abi.ENTITY_DTYPE or None (entity i == sphere i)
    triangles: np.ndarray = None  # abi.TRIANGLE_DTYPE or None
    focus_distance: float = None  # fixed focus for worlds the sphere-only auto-focus helper cannot walk
    placed: np.ndarray = None  # abi.PLACED_DTYPE or None: entities with the full Entity record (rtb_upload_placed_world)
    # image textures (rtb_upload_textures): uint8 [H, W, 3|4] arrays, one record per material, [n_triangles, 3, 2] vertex uvs
    images: list = None
    material_textures: np.ndarray = None  # abi.MATERIAL_TEXTURES_DTYPE or None
    triangle_uvs: np.ndarray = None


def generate_scene(scene_id, seed=700, target_count=0):
    """-> (spheres in scene order, materials, SceneInfo)."""
    info = abi.SceneInfo()
    rc = lib().rtbh_scene_generate(scene_id, seed, target_count, None, 0, None, 0, C.byref(info))
    if rc != 0:
        raise ValueError(f"rtbh_scene_generate failed: {rc}")
    spheres = np.zeros(info.sphere_count, dtype=abi.SPHERE_DTYPE)
    materials = np.zeros(info.material_count, dtype=abi.MATERIAL_DTYPE)
    rc = lib().rtbh_scene_generate(
        scene_id, seed, target_count, spheres.ctypes.data, len(spheres), materials.ctypes.data, len(materials), C.byref(info)
    )
    if rc != 0:
        raise ValueError(f"rtbh_scene_generate failed: {rc}")
    return spheres, materials, info


def build_bvh(spheres, max_depth):
    """RebuildBvh (Raytracer.cs:1306-1351): -> (BVH-ordered spheres, flattened nodes)."""
    spheres = np.ascontiguousarray(spheres, dtype=abi.SPHERE_DTYPE)
    n = len(spheres)
    out_spheres = np.zeros(n, dtype=abi.SPHERE_DTYPE)
    cap = max(1, 2 * n + 1)
    nodes = np.zeros(cap, dtype=abi.BVH_NODE_DTYPE)
    count = C.c_size_t(0)
    rc = lib().rtbh_build_bvh(
        spheres.ctypes.data if n else None, n, int(max_depth), out_spheres.ctypes.data if n else None, n,
        nodes.ctypes.data, cap, C.byref(count),
    )
    if rc != 0:
        raise ValueError(f"rtbh_build_bvh failed: {rc}")
    return out_spheres, nodes[: count.value].copy()


def _v3(x):
    return abi.f32x3(*[float(v) for v in x])


def make_view(origin, look_at, up, vertical_fov, aspect, aperture, focus_distance):
    """new View(...) (View.cs:16-36)."""
    v = abi.View()
    lib().rtbh_make_view(_v3(origin), _v3(look_at), _v3(up), vertical_fov, aspect, aperture, focus_distance, C.byref(v))
    return v


def hit_world(scene, origin, direction):
    """Raytracer.HitWorld: distance of the first hit along a ray, or None."""
    d = C.c_float(0)
    hit = lib().rtbh_hit_world(
        scene.nodes.ctypes.data, len(scene.nodes), scene.spheres.ctypes.data, len(scene.spheres),
        _v3(origin), _v3(direction), C.byref(d),
    )
    return d.value if hit else None


def view_for(scene, width, height, aperture=None, fallback_focus=1.0):
    """The camera block of ScheduleSample (Raytracer.cs:604-612) for a legacy-asset camera."""
    cam = abi.Camera.from_buffer_copy(scene.camera)
    if aperture is not None:
        cam.aperture = aperture
    if scene.focus_distance is not None:
        v = make_view(list(cam.position), list(cam.target), (0, 1, 0), cam.vertical_fov, float(width) / float(height),
                      cam.aperture, scene.focus_distance)
        return v, scene.focus_distance
    v = abi.View()
    focus = C.c_float(0)
    lib().rtbh_view_from_camera(
        C.byref(cam), float(width) / float(height), scene.nodes.ctypes.data, len(scene.nodes),
        scene.spheres.ctypes.data, len(scene.spheres), fallback_focus, C.byref(v), C.byref(focus),
    )
    return v, focus.value


def space_filling_series(length):
    out = (C.c_int32 * length)()
    rc = lib().rtbh_space_filling_series(length, out, length)
    if rc != 0:
        raise ValueError("rtbh_space_filling_series failed")
    return list(out)


def make_scene(name, max_bvh_depth=None, seed=700, target_count=0):
    """Named BASELINE scenes: 'three_spheres', 'final', 'stress'."""
    ids = {"three_spheres": abi.SCENE_THREE_SPHERES, "final": abi.SCENE_FINAL, "stress": abi.SCENE_STRESS}
    if name not in ids:
        raise ValueError(f"unknown scene {name!r}")
    if max_bvh_depth is None:
        max_bvh_depth = 0 if name == "three_spheres" else 16
    spheres, materials, info = generate_scene(ids[name], seed, target_count)
    bvh_spheres, nodes = build_bvh(spheres, max_bvh_depth)
    return Scene(
        spheres=bvh_spheres, materials=materials, nodes=nodes, camera=info.camera, environment=info.environment,
        info=info, max_bvh_depth=max_bvh_depth, name=name,
    )


def make_triangle(v1, v2, v3, material, normals=None):
    """new Triangle(...) (Triangle.cs:14-29): face normal, or the three vertex normals."""
    out = np.zeros(1, dtype=abi.TRIANGLE_DTYPE)
    ns = [None, None, None]
    keep = []
    if normals is not None:
        keep = [np.ascontiguousarray(n, dtype=np.float32) for n in normals]
        ns = [k.ctypes.data for k in keep]
    lib().rtbh_make_triangle(_v3(v1), _v3(v2), _v3(v3), ns[0], ns[1], ns[2], int(material), out.ctypes.data)
    return out[0]


def quat_from_to(a, b):
    """Unit quaternion (x, y, z, w) turning direction a into direction b (float64 arithmetic, rounded once)."""
    a = np.asarray(a, np.float64) / np.linalg.norm(a)
    b = np.asarray(b, np.float64) / np.linalg.norm(b)
    d = float(np.dot(a, b))
    if d < -1 + 1e-12:      # opposite: half a turn about any axis perpendicular to a
        axis = np.cross(a, (1.0, 0.0, 0.0)) if abs(a[0]) < 0.9 else np.cross(a, (0.0, 1.0, 0.0))
        axis /= np.linalg.norm(axis)
        return np.array([axis[0], axis[1], axis[2], 0.0], np.float32)
    q = np.array([*np.cross(a, b), 1.0 + d])
    return (q / np.linalg.norm(q)).astype(np.float32)


def quat_axis_angle(axis, degrees):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    h = np.radians(degrees) / 2
    return np.array([*(axis * np.sin(h)), np.cos(h)], np.float32)


def make_placed(entity_type, material, size, position, rotation=(0, 0, 0, 1), destination_offset=None, time_range=(0.0, 1.0)):
    """new Entity(type, content, new RigidTransform(rotation, position), material, moving, destinationOffset, timeRange)
    (Entity.cs:39-55) with its content: Sphere(radius) / Rect(size.xy) / Box(size.xyz)."""
    e = np.zeros(1, dtype=abi.PLACED_DTYPE)[0]
    e["type"], e["material"] = entity_type, material
    e["rotation"], e["position"] = rotation, position
    sz = np.zeros(3, np.float32)
    sz[: len(np.atleast_1d(size))] = np.atleast_1d(size)
    e["size"] = sz
    if destination_offset is not None:
        e["moving"], e["destination_offset"], e["time_range"] = 1, destination_offset, time_range
    return e


def add_mesh(vertices, indices, material, normals=None, uvs=None, rotation=(0, 0, 0, 1), position=(0, 0, 0), scale=1.0):
    """AddMeshRuntimeEntitiesJob: (triangles, per-triangle uvs [n, 3, 2]) of one mesh renderer — transform baked in,
    vertex normals rotated (normals=None: face normals)."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.uint16).reshape(-1)
    nrm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    tex = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
    rot = np.ascontiguousarray(rotation, dtype=np.float32)
    pos = np.ascontiguousarray(position, dtype=np.float32)
    n = len(idx) // 3
    tris = np.zeros(n, dtype=abi.TRIANGLE_DTYPE)
    out_uv = np.zeros((n, 3, 2), np.float32)
    count = C.c_size_t(0)
    rc = lib().rtbh_add_mesh(v.ctypes.data, None if nrm is None else nrm.ctypes.data, None if tex is None else tex.ctypes.data, len(v),
                             idx.ctypes.data, len(idx), rot.ctypes.data, pos.ctypes.data, float(scale), int(material),
                             tris.ctypes.data if n else None, out_uv.ctypes.data, n, C.byref(count))
    if rc != 0:
        raise ValueError(f"rtbh_add_mesh failed: {rc}")
    return tris[: count.value], out_uv[: count.value]


def build_world(spheres, triangles, materials, max_bvh_depth, camera, environment, focus_distance, name="world", placed=None):
    """RebuildWorld for a mixed world: per-entity bounds (CreateBvhBuildingEntitiesJob), the reference's BVH
    build + flatten order, and the BVH-ordered entity list the leaves index (bvhEntities).  Entities are
    listed spheres first, then triangles, then placed entities."""
    spheres = np.ascontiguousarray(spheres, dtype=abi.SPHERE_DTYPE)
    triangles = np.ascontiguousarray(triangles, dtype=abi.TRIANGLE_DTYPE)
    placed = np.ascontiguousarray(placed if placed is not None else [], dtype=abi.PLACED_DTYPE)
    n_st = len(spheres) + len(triangles)
    n = n_st + len(placed)
    bounds = np.zeros((n, 6), np.float32)
    for i in range(len(spheres)):
        lib().rtbh_sphere_bounds(spheres[i:i + 1].ctypes.data, bounds[i].ctypes.data)
    for i in range(len(triangles)):
        lib().rtbh_triangle_bounds(triangles[i:i + 1].ctypes.data, bounds[len(spheres) + i].ctypes.data)
    for i in range(len(placed)):
        lib().rtbh_placed_bounds(placed[i:i + 1].ctypes.data, bounds[n_st + i].ctypes.data)
    order = np.zeros(max(n, 1), np.uint32)
    cap = max(1, 2 * n + 1)
    nodes = np.zeros(cap, dtype=abi.BVH_NODE_DTYPE)
    count = C.c_size_t(0)
    rc = lib().rtbh_build_bvh_from_bounds(bounds.ctypes.data if n else None, n, int(max_bvh_depth), order.ctypes.data, n,
                                          nodes.ctypes.data, cap, C.byref(count))
    if rc != 0:
        raise ValueError(f"rtbh_build_bvh_from_bounds failed: {rc}")
    entities = np.zeros(n, dtype=abi.ENTITY_DTYPE)
    for k in range(n):
        e = int(order[k])
        if e < len(spheres):
            entities[k] = (abi.ENTITY_SPHERE, e)
        elif e < n_st:
            entities[k] = (abi.ENTITY_TRIANGLE, e - len(spheres))
        else:
            entities[k] = (int(placed[e - n_st]["type"]) | abi.ENTITY_PLACED, e - n_st)
    info = abi.SceneInfo()
    info.camera = camera
    info.environment = environment
    info.sphere_count = len(spheres)
    info.material_count = len(materials)
    return Scene(spheres=spheres, materials=np.ascontiguousarray(materials, dtype=abi.MATERIAL_DTYPE), nodes=nodes[: count.value].copy(),
                 camera=camera, environment=environment, info=info, max_bvh_depth=max_bvh_depth, name=name,
                 entities=entities, triangles=triangles, focus_distance=focus_distance,
                 placed=placed if len(placed) else None)


def _icosphere(center, radius, subdivisions):
    """Unit icosahedron subdivided `subdivisions` times: (vertices, faces); vertex normals = normalised offsets."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    normals = np.array(v, np.float32)
    verts = (np.array(center, np.float64) + radius * np.array(v)).astype(np.float32)
    return verts, normals, f


def _test_images():
    """Small deterministic images in the two formats the reference's host accepts (RGB24, RGBA32; Raytracer.cs:1214-1265)."""
    yy, xx = np.mgrid[0:32, 0:32]
    checker = np.zeros((32, 32, 3), np.uint8)
    on = ((xx // 4 + yy // 4) % 2) == 0
    checker[on] = (230, 220, 200)
    checker[~on] = (40, 60, 90)
    checker[:, :, 0] = np.where(xx > 24, 255, checker[:, :, 0])
    rgba = np.zeros((16, 24, 4), np.uint8)                     # base colour map whose alpha is the smoothness
    y2, x2 = np.mgrid[0:16, 0:24]
    rgba[..., 0] = 120 + 5 * x2
    rgba[..., 1] = 250 - 9 * y2
    rgba[..., 2] = 90 + ((x2 * 7 + y2 * 13) % 160)
    rgba[..., 3] = np.where((x2 + y2) % 3 == 0, 255, 60 + 8 * y2)
    metal = np.zeros((8, 8, 3), np.uint8)
    metal[..., 0] = np.where(np.mgrid[0:8, 0:8][1] % 2 == 0, 255, 0)
    glow = np.zeros((4, 4, 3), np.uint8)
    glow[..., 0], glow[..., 1], glow[..., 2] = 255, 180, 60
    glow[1:3, 1:3] = (20, 60, 255)
    return [checker, rgba, metal, glow]


def make_mesh_scene(max_bvh_depth=16, subdivisions=1, emissive=False, textured=False):
    """A small mixed world in the shape of what the reference's host produces at HEAD (mesh triangles,
    Raytracer.cs:1185-1304) plus sphere entities: a two-triangle ground quad (face normals), a smooth-shaded
    icosphere (vertex normals, glossy metal), a flat-shaded tetrahedron (Lambertian), a hollow glass sphere
    (negative-radius inner sphere) and a diffuse sphere; gradient sky.  emissive=True: no sky (SkyType.None,
    Environment.cs:5-10) and an emissive panel overhead (Material.Emit, Material.cs:175-179) as the only light."""
    def mat(mtype, albedo, gloss=0.0, metallic=0.0, ior=1.5, emission=(0, 0, 0)):
        m = np.zeros(1, dtype=abi.MATERIAL_DTYPE)[0]
        m["type"], m["albedo"], m["emission"] = mtype, albedo, emission
        m["glossiness"], m["metallic"], m["index_of_refraction"] = gloss, metallic, ior
        return m

    materials = np.array([
        mat(abi.MATERIAL_STANDARD, (0.5, 0.5, 0.5)),                               # 0 ground
        mat(abi.MATERIAL_STANDARD, (0.8, 0.6, 0.2), gloss=0.8, metallic=1.0),      # 1 icosphere
        mat(abi.MATERIAL_STANDARD, (0.7, 0.15, 0.1)),                              # 2 tetrahedron
        mat(abi.MATERIAL_DIELECTRIC, (1, 1, 1), gloss=1.0, ior=1.5),               # 3 glass
        mat(abi.MATERIAL_STANDARD, (0.1, 0.2, 0.5)),                               # 4 small sphere
        mat(abi.MATERIAL_STANDARD, (0.9, 0.9, 0.9), gloss=0.4, metallic=0.3),      # 5 glossy part-metal wedge
        mat(abi.MATERIAL_STANDARD, (0.3, 0.3, 0.3), emission=(9.0, 8.0, 6.0)),     # 6 light panel
    ], dtype=abi.MATERIAL_DTYPE)
    tris = []
    g = 12.0
    q = [(-g, 0, -g), (g, 0, -g), (g, 0, g), (-g, 0, g)]
    tris += [make_triangle(q[0], q[2], q[1], 0), make_triangle(q[0], q[3], q[2], 0)]
    verts, normals, faces = _icosphere((0.0, 1.0, 0.0), 1.0, subdivisions)
    for a, b, c in faces:
        tris.append(make_triangle(verts[a], verts[c], verts[b], 1, normals=(normals[a], normals[c], normals[b])))
    p = np.array([(2.2, 0.0, -0.8), (3.6, 0.0, -0.6), (2.9, 0.0, 0.6), (2.9, 1.3, -0.2)], np.float32)
    for a, b, c in ((0, 1, 3), (1, 2, 3), (2, 0, 3), (0, 2, 1)):
        tris.append(make_triangle(p[a], p[c], p[b], 2))
    w = np.array([(0.3, 0.0, -3.2), (1.5, 0.0, -2.7), (0.8, 1.1, -3.0)], np.float32)
    tris.append(make_triangle(w[0], w[1], w[2], 5))          # a lone two-sided wedge (hit from both faces)
    if emissive:
        l = [(-2.5, 4.0, -2.5), (2.5, 4.0, -2.5), (2.5, 4.0, 2.5), (-2.5, 4.0, 2.5)]
        tris += [make_triangle(l[0], l[1], l[2], 6), make_triangle(l[0], l[2], l[3], 6)]
    spheres = np.zeros(3, dtype=abi.SPHERE_DTYPE)
    spheres[0] = ((-1.9, 0.7, -1.2), 0.7, 3, (0, 0, 0))
    spheres[1] = ((-1.9, 0.7, -1.2), -0.62, 3, (0, 0, 0))
    spheres[2] = ((1.2, 0.35, -1.9), 0.35, 4, (0, 0, 0))
    cam = abi.Camera()
    cam.position[:] = (5.0, 2.6, -6.5)
    cam.target[:] = (0.0, 0.8, 0.0)
    cam.aperture = 0.0
    cam.vertical_fov = 32.0
    env = abi.Environment()
    env.sky_type = abi.SKY_NONE if emissive else abi.SKY_GRADIENT
    env.sky_bottom_color[:] = (1.0, 1.0, 1.0)
    env.sky_top_color[:] = (0.5, 0.7, 1.0)
    focus = float(np.linalg.norm(np.array(cam.position[:]) - np.array(cam.target[:])))
    scene = build_world(spheres, np.array(tris, dtype=abi.TRIANGLE_DTYPE), materials, max_bvh_depth, cam, env, focus, name="mesh")
    if textured:
        # what the host builds for Unity materials with maps (Raytracer.cs:1210-1267): a base-colour map on the ground
        # (RGB24, uv reaching exactly 1 at the far edges), an RGBA32 base-colour map on the icosphere whose alpha is its
        # smoothness plus a metallic map, an emissive map on the light panel; sphere entities have TexCoords = 0
        t = scene.triangles
        v = np.stack([t["v0"], t["v0"] + t["edge1"], t["v0"] + t["edge2"]], axis=1).astype(np.float64)   # ctor order v1, v2, v3
        uv = np.zeros((len(t), 3, 2), np.float32)
        m = t["material"]
        uv[m == 0] = ((v[m == 0][:, :, [0, 2]] + g) / (2 * g)).astype(np.float32)
        c = v[m == 1] - np.array([0.0, 1.0, 0.0])
        uv[m == 1, :, 0] = (np.arctan2(c[..., 2], c[..., 0]) / (2 * np.pi) + 0.5).astype(np.float32)
        uv[m == 1, :, 1] = (np.arccos(np.clip(c[..., 1], -1, 1)) / np.pi).astype(np.float32)
        uv[m == 2] = (v[m == 2][:, :, [0, 1]] % 1.0).astype(np.float32)
        uv[m == 6] = ((v[m == 6][:, :, [0, 2]] + 2.5) / 5.0).astype(np.float32)
        mt = np.zeros(len(materials), dtype=abi.MATERIAL_TEXTURES_DTYPE)
        for k in ("albedo_image", "emission_image", "glossiness_image", "metallic_image"):
            mt[k] = -1
        mt[0]["albedo_image"] = 0
        mt[1]["albedo_image"], mt[1]["glossiness_image"], mt[1]["glossiness_channel"], mt[1]["metallic_image"] = 1, 1, 3, 2
        mt[2]["albedo_image"] = 1
        mt[3]["glossiness_image"], mt[3]["glossiness_channel"] = 1, 3
        mt[4]["albedo_image"] = 0
        mt[6]["emission_image"] = 3
        scene.images, scene.material_textures, scene.triangle_uvs = _test_images(), mt, uv
    return scene


def _material(mtype, albedo, gloss=0.0, metallic=0.0, ior=1.5, emission=(0, 0, 0)):
    m = np.zeros(1, dtype=abi.MATERIAL_DTYPE)[0]
    m["type"], m["albedo"], m["emission"] = mtype, albedo, emission
    m["glossiness"], m["metallic"], m["index_of_refraction"] = gloss, metallic, ior
    return m


def make_cornell_scene(max_bvh_depth=16, moving=True, fog=False):
    """A Cornell box in the reference's entity vocabulary (the kind of world its Rect / Box entities exist for):
    five Rect walls and an emissive Rect light (Rect.cs: an XY rectangle hit from +Z only, turned into place by the
    entity's rotation), two Box entities turned about Y, a glass sphere (plain sphere entity) and — moving=True — a
    sphere that moves during the exposure (Entity.TransformAtTime) plus a moving, rotated box.  The front is open and
    there is no sky: the light panel is the only emitter (Material.Emit), a path ends when it leaves through the front."""
    S = 5.55
    materials = np.array([
        _material(abi.MATERIAL_STANDARD, (0.73, 0.73, 0.73)),                          # 0 white
        _material(abi.MATERIAL_STANDARD, (0.65, 0.05, 0.05)),                          # 1 red
        _material(abi.MATERIAL_STANDARD, (0.12, 0.45, 0.15)),                          # 2 green
        _material(abi.MATERIAL_STANDARD, (0.0, 0.0, 0.0), emission=(15.0, 15.0, 15.0)),  # 3 light
        _material(abi.MATERIAL_DIELECTRIC, (1, 1, 1), gloss=1.0, ior=1.5),             # 4 glass
        _material(abi.MATERIAL_STANDARD, (0.8, 0.8, 0.9), gloss=0.9, metallic=1.0),    # 5 polished metal (tall box)
        _material(abi.MATERIAL_STANDARD, (0.2, 0.3, 0.8)),                             # 6 blue (moving sphere)
        _material(abi.MATERIAL_PROBABILISTIC_VOLUME, (0.05, 0.05, 0.05), ior=0.9),     # 7 dark smoke (density in the parameter field)
        _material(abi.MATERIAL_PROBABILISTIC_VOLUME, (0.95, 0.95, 0.95), ior=0.6),     # 8 white fog
        _material(abi.MATERIAL_PROBABILISTIC_VOLUME, (0.3, 0.5, 0.9), ior=2.5),        # 9 dense blue medium inside the glass ball
    ], dtype=abi.MATERIAL_DTYPE)
    if not fog:
        materials = materials[:7]
    z = (0.0, 0.0, 1.0)
    h = S / 2
    placed = [
        make_placed(abi.ENTITY_RECT, 0, (S, S), (h, 0.0, h), quat_from_to(z, (0, 1, 0))),        # floor
        make_placed(abi.ENTITY_RECT, 0, (S, S), (h, S, h), quat_from_to(z, (0, -1, 0))),         # ceiling
        make_placed(abi.ENTITY_RECT, 0, (S, S), (h, h, S), quat_from_to(z, (0, 0, -1))),         # back wall
        make_placed(abi.ENTITY_RECT, 1, (S, S), (0.0, h, h), quat_from_to(z, (1, 0, 0))),        # red wall
        make_placed(abi.ENTITY_RECT, 2, (S, S), (S, h, h), quat_from_to(z, (-1, 0, 0))),         # green wall
        make_placed(abi.ENTITY_RECT, 3, (1.3, 1.05), (h, S - 0.01, h), quat_from_to(z, (0, -1, 0))),   # light
        make_placed(abi.ENTITY_BOX, 5, (1.65, 3.3, 1.65), (3.6, 1.65, 3.5), quat_axis_angle((0, 1, 0), 15.0)),
        make_placed(abi.ENTITY_BOX, 0, (1.65, 1.65, 1.65), (1.8, 0.825, 1.7), quat_axis_angle((0, 1, 0), -18.0)),
    ]
    if moving:
        placed.append(make_placed(abi.ENTITY_SPHERE, 6, 0.45, (4.4, 0.45, 1.2), quat_axis_angle((1, 2, 3), 40.0),
                                  destination_offset=(0.0, 0.5, 0.3), time_range=(0.0, 1.0)))
        placed.append(make_placed(abi.ENTITY_BOX, 1, (0.5, 0.5, 0.5), (0.9, 3.6, 3.9), quat_axis_angle((1, 1, 0), 30.0),
                                  destination_offset=(0.4, -0.3, 0.0), time_range=(0.25, 0.75)))
    spheres = np.zeros(1, dtype=abi.SPHERE_DTYPE)
    spheres[0] = ((1.8, 1.65 + 0.6, 1.7), 0.6, 4, (0, 0, 0))
    if fog:
        # participating media (Material.ProbabilisticHit): balls of fog and smoke, a denser medium inside the glass ball
        # (a volume sphere just inside it), and the moving sphere becomes a drifting ball of fog
        # (a Box medium is inert in the reference — Box.Hit's normal always faces the ray, HitTests.cs:108 — so the short box
        # only exercises that path; the visible media are spheres)
        placed[7]["material"] = 8
        spheres = np.concatenate([spheres, np.zeros(3, dtype=abi.SPHERE_DTYPE)])
        spheres[1] = ((1.8, 1.65 + 0.6, 1.7), 0.55, 9, (0, 0, 0))
        spheres[2] = ((3.9, 1.3, 1.6), 1.2, 8, (0, 0, 0))        # a ball of white fog in front of the tall box
        spheres[3] = ((1.2, 3.6, 3.2), 1.0, 7, (0, 0, 0))        # dark smoke under the ceiling, overlapping the moving box
        if moving:
            placed[8]["material"] = 8
    cam = abi.Camera()
    cam.position[:] = (h, h, -8.0)
    cam.target[:] = (h, h, 0.0)
    cam.aperture = 0.0
    cam.vertical_fov = 40.0
    env = abi.Environment()
    env.sky_type = abi.SKY_NONE
    return build_world(spheres, [], materials, max_bvh_depth, cam, env, 8.0 + h, name="cornell",
                       placed=np.array(placed, dtype=abi.PLACED_DTYPE))


def make_random_placed_scene(seed, count=24, max_bvh_depth=8, media=0):
    """A random world of every entity kind (plain spheres, rotated / moving spheres, Rects, Boxes, a few triangles) with
    random Standard / Dielectric / emissive materials, inside a gradient sky: the fuzz input of the parity tests.
    media = k turns k of the eight materials into ProbabilisticVolume media of random density (whatever entities wear
    them: overlapping, nested, moving, non-convex ones too), without changing anything else of the world."""
    rng = np.random.default_rng(seed)
    mats = []
    for _ in range(8):
        u = rng.random()
        if u < 0.5:
            mats.append(_material(abi.MATERIAL_STANDARD, rng.random(3), gloss=float(rng.choice([0.0, 0.0, 0.6, 1.0])),
                                  metallic=float(rng.choice([0.0, 0.0, 0.5, 1.0]))))
        elif u < 0.75:
            mats.append(_material(abi.MATERIAL_DIELECTRIC, (1, 1, 1), gloss=float(rng.choice([1.0, 0.8])), ior=1.5))
        else:
            mats.append(_material(abi.MATERIAL_STANDARD, rng.random(3) * 0.5, emission=rng.random(3) * 4))
    if media:
        mrng = np.random.default_rng(seed + 7919)
        for k in mrng.choice(len(mats), size=media, replace=False):
            mats[int(k)] = _material(abi.MATERIAL_PROBABILISTIC_VOLUME, mrng.random(3), ior=float(0.1 + 3.0 * mrng.random()))
    materials = np.array(mats, dtype=abi.MATERIAL_DTYPE)

    def rand_quat():
        q = rng.normal(size=4)
        return (q / np.linalg.norm(q)).astype(np.float32)

    spheres, placed, tris = [], [], []
    placed.append(make_placed(abi.ENTITY_RECT, 0, (14.0, 14.0), (0.0, -0.05, 0.0), quat_from_to((0, 0, 1), (0, 1, 0))))   # a floor
    for _ in range(count):
        pos = (rng.random(3) * (6, 2.5, 6) - (3, 0, 3)).astype(np.float32)
        m = int(rng.integers(len(mats)))
        kind = rng.integers(6)
        motion = dict(destination_offset=(rng.random(3) - 0.5).astype(np.float32) * 2,
                      time_range=(float(rng.random() * 0.5), float(0.5 + rng.random() * 0.5))) if rng.random() < 0.3 else {}
        if kind == 0:
            spheres.append((tuple(pos), float(0.3 + rng.random() * 0.9) * (1 if rng.random() < 0.9 else -1), m, (0, 0, 0)))
        elif kind == 1:
            placed.append(make_placed(abi.ENTITY_SPHERE, m, float(0.3 + rng.random() * 0.9), pos, rand_quat(), **motion))
        elif kind in (2, 3):
            rot = rand_quat() if rng.random() < 0.7 else quat_from_to((0, 0, 1), [(0, 1, 0), (1, 0, 0), (0, 0, -1)][int(rng.integers(3))])
            placed.append(make_placed(abi.ENTITY_RECT, m, rng.random(2) * 3 + 0.3, pos, rot, **motion))
        elif kind == 4:
            rot = rand_quat() if rng.random() < 0.7 else (0, 0, 0, 1)
            placed.append(make_placed(abi.ENTITY_BOX, m, rng.random(3) * 2.0 + 0.3, pos, rot, **motion))
        else:
            a = pos + (rng.random(3) - 0.5) * 3
            tris.append(make_triangle(pos, a, pos + (rng.random(3) - 0.5) * 4, m))
    cam = abi.Camera()
    cam.position[:] = (0.3, 1.6, -9.0)
    cam.target[:] = (0.0, 1.2, 0.0)
    cam.aperture = 0.0
    cam.vertical_fov = 45.0
    env = abi.Environment()
    env.sky_type = abi.SKY_GRADIENT
    env.sky_bottom_color[:] = (1.0, 1.0, 1.0)
    env.sky_top_color[:] = (0.5, 0.7, 1.0)
    return build_world(np.array(spheres, dtype=abi.SPHERE_DTYPE) if spheres else np.zeros(0, abi.SPHERE_DTYPE),
                       np.array(tris, dtype=abi.TRIANGLE_DTYPE) if tris else np.zeros(0, abi.TRIANGLE_DTYPE), materials, max_bvh_depth,
                       cam, env, 9.0, name=f"random{seed}", placed=np.array(placed, dtype=abi.PLACED_DTYPE) if placed else None)


def make_params(scene, width, height, spp, trace_depth, seed=1, aperture=None, jitter=True,
                slice_offset=0, slice_divider=1, row_begin=0, row_end=0, spp_max=None):
    """Fills the uniform fields of the job the way ScheduleSample does (Raytracer.cs:671-712)."""
    p = abi.BatchParams()
    p.size[0], p.size[1] = float(width), float(height)
    p.slice_offset, p.slice_divider = slice_offset, slice_divider
    p.seed = seed
    p.view, _ = view_for(scene, width, height, aperture)
    p.environment = scene.environment
    p.sample_count_range[0] = spp
    p.sample_count_range[1] = spp if spp_max is None else spp_max
    p.trace_depth = trace_depth
    p.sub_pixel_jitter = 1 if jitter else 0
    p.sample_count_weight_extrema[0] = 0.0
    p.sample_count_weight_extrema[1] = 0.0
    p.row_begin, p.row_end = row_begin, row_end
    return p
