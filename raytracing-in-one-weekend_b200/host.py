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
