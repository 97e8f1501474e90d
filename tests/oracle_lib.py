"""Loader for the CPU oracle (oracle/liboracle*.so).  TEST-SIDE ONLY: the product package
never imports this module or anything under oracle/."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

rtb = importlib.import_module("raytracing-in-one-weekend_b200")
abi = rtb.abi

NOISE_XORSHIFT = 0  # the reference's Unity.Mathematics.Random stream
NOISE_PHILOX = 1  # the slot-keyed Philox stream the GPU uses

_libs = {}


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        path = os.path.join(ORACLE_DIR, name)
        src = os.path.join(ORACLE_DIR, "oracle.cpp")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            build_oracle()
        L = C.CDLL(path)
        L.oracle_sample_batch.argtypes = [
            C.POINTER(abi.BatchParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
            C.POINTER(abi.BatchBuffers), C.c_int, C.c_int, C.c_int64, C.c_int64,
        ]
        L.oracle_sample_batch.restype = C.c_int
        L.oracle_sample_batch_world.argtypes = [
            C.POINTER(abi.BatchParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_size_t, C.POINTER(abi.BatchBuffers), C.c_int, C.c_int, C.c_int64, C.c_int64,
        ]
        L.oracle_sample_batch_world.restype = C.c_int
        L.oracle_sample_batch_placed.argtypes = [
            C.POINTER(abi.BatchParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(abi.BatchBuffers), C.c_int, C.c_int, C.c_int64, C.c_int64,
        ]
        L.oracle_sample_batch_placed.restype = C.c_int
        L.oracle_placed_hit.argtypes = [C.c_void_p, abi.f32x3, abi.f32x3, C.c_float, C.POINTER(C.c_float), abi.f32x3, abi.f32x3]
        L.oracle_placed_hit.restype = C.c_int
        L.oracle_set_textures.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.oracle_set_textures.restype = None
        L.oracle_set_sky_cubemap.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_sky_cubemap.restype = None
        L.oracle_cubemap_sample.argtypes = [abi.f32x3, abi.f32x3]
        L.oracle_cubemap_sample.restype = None
        L.oracle_triangle_hit.argtypes = [C.c_void_p, abi.f32x3, abi.f32x3, C.POINTER(C.c_float), abi.f32x3, abi.f32x3]
        L.oracle_triangle_hit.restype = C.c_int
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_unity_random.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_sphere_hit.argtypes = [abi.f32x3, C.c_float, abi.f32x3, abi.f32x3, C.POINTER(C.c_float), abi.f32x3, abi.f32x3]
        L.oracle_sphere_hit.restype = C.c_int
        L.oracle_aabb_hit.argtypes = [abi.f32x3, abi.f32x3, abi.f32x3, abi.f32x3]
        L.oracle_aabb_hit.restype = C.c_int
        L.oracle_schlick.argtypes = [C.c_float, C.c_float]
        L.oracle_schlick.restype = C.c_float
        L.oracle_roughness_to_alpha.argtypes = [C.c_float]
        L.oracle_roughness_to_alpha.restype = C.c_float
        L.oracle_lambda.argtypes = [abi.f32x3, abi.f32x3, C.c_float]
        L.oracle_lambda.restype = C.c_float
        L.oracle_refract.argtypes = [abi.f32x3, abi.f32x3, C.c_float, abi.f32x3]
        L.oracle_refract.restype = C.c_int
        L.oracle_get_ray.argtypes = [
            C.POINTER(abi.View), C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, abi.f32x3, abi.f32x3
        ]
        L.oracle_umath_sincos.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_umath_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_umath_pow.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_int]
        L.oracle_umath_vec.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_umath_vec.restype = C.c_int
        _libs[name] = L
    return _libs[name]


class Buffers:
    """The accumulation buffers of one batch as numpy arrays (SampleBatchJob.cs:41-51)."""

    def __init__(self, width, height, diagnostics=True):
        n = width * height
        self.width, self.height = width, height
        self.in_color = np.zeros((n, 4), np.float32)
        self.in_weight = np.zeros(n, np.float32)
        self.in_normal = np.zeros((n, 3), np.float32)
        self.in_albedo = np.zeros((n, 3), np.float32)
        self.out_color = np.zeros((n, 4), np.float32)
        self.out_weight = np.zeros(n, np.float32)
        self.out_normal = np.zeros((n, 3), np.float32)
        self.out_albedo = np.zeros((n, 3), np.float32)
        self.diagnostics = np.zeros(n, abi.DIAGNOSTICS_DTYPE) if diagnostics else None

    def as_struct(self):
        b = abi.BatchBuffers()
        b.in_color = self.in_color.ctypes.data
        b.in_sample_count_weight = self.in_weight.ctypes.data
        b.in_normal = self.in_normal.ctypes.data
        b.in_albedo = self.in_albedo.ctypes.data
        b.out_color = self.out_color.ctypes.data
        b.out_sample_count_weight = self.out_weight.ctypes.data
        b.out_normal = self.out_normal.ctypes.data
        b.out_albedo = self.out_albedo.ctypes.data
        b.out_diagnostics = self.diagnostics.ctypes.data if self.diagnostics is not None else None
        return b

    def swap(self):
        """Ping-pong: accumulation := output (Raytracer.cs:798-802)."""
        self.in_color, self.out_color = self.out_color, self.in_color
        self.in_weight, self.out_weight = self.out_weight, self.in_weight
        self.in_normal, self.out_normal = self.out_normal, self.in_normal
        self.in_albedo, self.out_albedo = self.out_albedo, self.in_albedo

    def rgb(self):
        """CombineJob's per-pixel colour (CombineJob.cs:34-54): xyz / (int)w, 0 when no samples."""
        n = self.out_color[:, 3].astype(np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            rgb = self.out_color[:, :3] / np.maximum(n, 1)[:, None].astype(np.float32)
        rgb[n == 0] = 0
        return rgb.reshape(self.height, self.width, 3)


_sky_keepalive = {}


def set_sky_cubemap(faces, fast=False):
    """Environment.SkyCubemap for subsequent oracle batches: [6, H, W, 4] float16, or None."""
    L = lib(fast)
    if faces is None:
        _sky_keepalive.pop(fast, None)
        L.oracle_set_sky_cubemap(None, 0, 0)
        return
    f = np.ascontiguousarray(faces)
    if f.dtype == np.float16:
        f = f.view(np.uint16)
    _sky_keepalive[fast] = f
    L.oracle_set_sky_cubemap(f.ctypes.data, f.shape[2], f.shape[1])


_tex_keepalive = {}


def set_textures(scene, fast=False):
    """Image textures of `scene` (images, material_textures, triangle_uvs) for subsequent oracle batches, or none."""
    L = lib(fast)
    if scene is None or getattr(scene, "material_textures", None) is None:
        _tex_keepalive.pop(fast, None)
        L.oracle_set_textures(None, 0, None, 0, None, 0)
        return
    imgs, keep = rtb.plugin.image_structs(scene.images)
    mt = np.ascontiguousarray(scene.material_textures, dtype=abi.MATERIAL_TEXTURES_DTYPE)
    uv = None if scene.triangle_uvs is None else np.ascontiguousarray(scene.triangle_uvs, dtype=np.float32)
    _tex_keepalive[fast] = (imgs, keep, mt, uv)
    L.oracle_set_textures(C.addressof(imgs), len(keep), mt.ctypes.data, len(mt), None if uv is None else uv.ctypes.data,
                          0 if uv is None else uv.size // 6)


def sample_batch(scene, params, buffers, noise=NOISE_PHILOX, threads=None, fast=False, index_range=(0, 0)):
    threads = threads or os.cpu_count() or 1
    b = buffers.as_struct()
    set_textures(scene, fast)       # a scene's image textures travel with it (none: cleared)
    if getattr(scene, "placed", None) is not None and len(scene.placed):
        tris, sph = scene.triangles, scene.spheres
        rc = lib(fast).oracle_sample_batch_placed(
            C.byref(params), scene.entities.ctypes.data, len(scene.entities), sph.ctypes.data if len(sph) else None, len(sph),
            tris.ctypes.data if len(tris) else None, len(tris), scene.placed.ctypes.data, len(scene.placed),
            scene.materials.ctypes.data, len(scene.materials), scene.nodes.ctypes.data, len(scene.nodes), C.byref(b), noise, threads,
            index_range[0], index_range[1],
        )
        if rc != 0:
            raise RuntimeError(f"oracle_sample_batch_placed failed: {rc}")
        return buffers
    if getattr(scene, "entities", None) is not None:
        tris = scene.triangles
        rc = lib(fast).oracle_sample_batch_world(
            C.byref(params), scene.entities.ctypes.data, len(scene.entities), scene.spheres.ctypes.data, len(scene.spheres),
            tris.ctypes.data if len(tris) else None, len(tris), scene.materials.ctypes.data, len(scene.materials),
            scene.nodes.ctypes.data, len(scene.nodes), C.byref(b), noise, threads, index_range[0], index_range[1],
        )
        if rc != 0:
            raise RuntimeError(f"oracle_sample_batch_world failed: {rc}")
        return buffers
    rc = lib(fast).oracle_sample_batch(
        C.byref(params), scene.spheres.ctypes.data, len(scene.spheres), scene.materials.ctypes.data, len(scene.materials),
        scene.nodes.ctypes.data, len(scene.nodes), C.byref(b), noise, threads, index_range[0], index_range[1],
    )
    if rc != 0:
        raise RuntimeError(f"oracle_sample_batch failed: {rc}")
    return buffers


UMATH_VEC_SIZES = {0: (6, 3), 1: (7, 3), 2: (7, 7), 3: (10, 3), 4: (7, 3), 5: (3, 3), 6: (6, 3), 7: (12, 3), 8: (6, 1), 9: (3, 8)}


def umath_vec(op, records):
    """include/rtb/umath.h's vector / quaternion functions on float32 records [n, in_size] -> [n, out_size]."""
    a = np.ascontiguousarray(records, dtype=np.float32)
    n_in, n_out = UMATH_VEC_SIZES[op]
    assert a.ndim == 2 and a.shape[1] == n_in
    out = np.zeros((len(a), n_out), np.float32)
    assert lib().oracle_umath_vec(op, a.ctypes.data, out.ctypes.data, len(a)) == 0
    return out
