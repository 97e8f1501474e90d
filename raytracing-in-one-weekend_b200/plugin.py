"""ctypes binding of librtb.so (include/rtb.h) — the sm_100a plugin.

This is the Python twin of the P/Invoke class a Unity maintainer would add
(bindings/B200PathTracerApi.cs).  There is NO CPU fallback: loading
fails loudly when the library has not been built, and `Context()` fails when no sm_100
device is present.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi
from . import build as _build

EXPORTS = [
    "rtb_abi_version", "rtb_create", "rtb_destroy", "rtb_last_error", "rtb_set_log_callback",
    "rtb_upload_scene", "rtb_upload_world", "rtb_upload_placed_world", "rtb_upload_textures", "rtb_upload_sky_cubemap", "rtb_describe_scene", "rtb_retree_bvh", "rtb_sample_batch", "rtb_sample_batch_device",
    "rtb_register_host_buffer", "rtb_unregister_host_buffer",
    "rtb_combine_device", "rtb_finalize_device", "rtb_reduce_metrics_device",
    "rtb_get_counters", "rtb_set_option", "rtb_last_kernel_ms", "rtb_last_batch_in_place", "rtb_measure_fp32_peak",
    "rtb_multi_create", "rtb_multi_destroy", "rtb_multi_device_count", "rtb_multi_context", "rtb_multi_last_error", "rtb_multi_set_option",
    "rtb_multi_upload_scene", "rtb_multi_upload_placed_world", "rtb_multi_upload_textures", "rtb_multi_upload_sky_cubemap",
    "rtb_multi_register_host_buffer", "rtb_multi_unregister_host_buffer", "rtb_multi_sample_batch", "rtb_multi_sample_batch_device",
    "rtb_multi_get_tiles", "rtb_balance_rows",
    "rtb_device_alloc", "rtb_device_free", "rtb_ipc_export", "rtb_ipc_open", "rtb_ipc_close",
]
_NOT_INT = ("rtb_last_error", "rtb_multi_last_error", "rtb_multi_context")

_lib = None


class RtbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rtb error {code}: {message}")
        self.code = code


def lib():
    """Loads lib/librtb.so (never builds it implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is None:
        path = os.environ.get("RTB_PLUGIN_LIB") or _build.plugin_lib_path()   # override: experiment builds only
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback"
            )
        L = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        L.rtb_abi_version.restype = C.c_int
        L.rtb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.rtb_destroy.argtypes = [vp]
        L.rtb_last_error.argtypes = [vp]
        L.rtb_last_error.restype = C.c_char_p
        L.rtb_set_log_callback.argtypes = [vp, vp, vp]
        L.rtb_upload_scene.argtypes = [vp, vp, sz, vp, sz, vp, sz]
        L.rtb_upload_world.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
        L.rtb_upload_placed_world.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
        L.rtb_upload_textures.argtypes = [vp, vp, sz, vp, sz, vp, sz]
        L.rtb_upload_sky_cubemap.argtypes = [vp, vp, C.c_int, C.c_int]
        L.rtb_describe_scene.argtypes = [vp, sz, vp, sz, vp, sz, C.c_int, C.POINTER(abi.SceneLayout)]
        L.rtb_retree_bvh.argtypes = [vp, sz, vp, sz, C.POINTER(C.c_size_t)]
        L.rtb_sample_batch.argtypes = [vp, C.POINTER(abi.BatchParams), C.POINTER(abi.BatchBuffers), vp]
        L.rtb_sample_batch_device.argtypes = [vp, C.POINTER(abi.BatchParams), C.POINTER(abi.BatchBuffers), vp]
        L.rtb_register_host_buffer.argtypes = [vp, vp, sz]
        L.rtb_unregister_host_buffer.argtypes = [vp, vp]
        L.rtb_combine_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.rtb_finalize_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
        L.rtb_reduce_metrics_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.POINTER(abi.Metrics), vp]
        L.rtb_get_counters.argtypes = [vp, C.POINTER(abi.Counters)]
        L.rtb_set_option.argtypes = [vp, C.c_int, C.c_int64]
        L.rtb_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.rtb_last_batch_in_place.argtypes = [vp, C.POINTER(C.c_int)]
        L.rtb_measure_fp32_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.rtb_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        L.rtb_multi_destroy.argtypes = [vp]
        L.rtb_multi_device_count.argtypes = [vp]
        L.rtb_multi_context.argtypes = [vp, C.c_int]
        L.rtb_multi_context.restype = vp
        L.rtb_multi_last_error.argtypes = [vp]
        L.rtb_multi_last_error.restype = C.c_char_p
        L.rtb_multi_set_option.argtypes = [vp, C.c_int, C.c_int64]
        L.rtb_multi_upload_scene.argtypes = [vp, vp, sz, vp, sz, vp, sz]
        L.rtb_multi_upload_placed_world.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
        L.rtb_multi_upload_textures.argtypes = [vp, vp, sz, vp, sz, vp, sz]
        L.rtb_multi_upload_sky_cubemap.argtypes = [vp, vp, C.c_int, C.c_int]
        L.rtb_multi_register_host_buffer.argtypes = [vp, vp, sz]
        L.rtb_multi_unregister_host_buffer.argtypes = [vp, vp]
        L.rtb_multi_sample_batch.argtypes = [vp, C.POINTER(abi.BatchParams), C.POINTER(abi.BatchBuffers), vp]
        L.rtb_multi_sample_batch_device.argtypes = [vp, C.POINTER(abi.BatchParams), C.POINTER(abi.BatchBuffers), C.c_int, vp]
        L.rtb_multi_get_tiles.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_float)]
        L.rtb_balance_rows.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.rtb_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
        L.rtb_device_free.argtypes = [vp, vp]
        L.rtb_ipc_export.argtypes = [vp, vp, vp]
        L.rtb_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
        L.rtb_ipc_close.argtypes = [vp, vp]
        for name in EXPORTS:
            if name not in _NOT_INT:
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


class HostBuffers:
    """The eight accumulation arrays of one batch (SampleBatchJob.cs:41-51) + diagnostics, as
    numpy arrays in the reference's layout (index = row * W + col, row 0 = bottom)."""

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

    def arrays(self):
        a = [self.in_color, self.in_weight, self.in_normal, self.in_albedo,
             self.out_color, self.out_weight, self.out_normal, self.out_albedo]
        if self.diagnostics is not None:
            a.append(self.diagnostics)
        return a

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
        """accumulation := output (Raytracer.cs:798-802)."""
        self.in_color, self.out_color = self.out_color, self.in_color
        self.in_weight, self.out_weight = self.out_weight, self.in_weight
        self.in_normal, self.out_normal = self.out_normal, self.in_normal
        self.in_albedo, self.out_albedo = self.out_albedo, self.in_albedo

    def rgb(self):
        """Per-pixel colour as CombineJob defines it (CombineJob.cs:34-54)."""
        n = self.out_color[:, 3].astype(np.int32)
        rgb = self.out_color[:, :3] / np.maximum(n, 1)[:, None].astype(np.float32)
        rgb[n == 0] = 0
        return rgb.reshape(self.height, self.width, 3)


def device_buffers_struct(in_color, in_weight, in_normal, in_albedo, out_color, out_weight, out_normal, out_albedo,
                          diagnostics=None):
    """rtb_batch_buffers from objects exposing `.data_ptr()` (torch CUDA tensors) or raw ints."""
    def ptr(x):
        if x is None:
            return None
        return x.data_ptr() if hasattr(x, "data_ptr") else int(x)

    b = abi.BatchBuffers()
    b.in_color, b.in_sample_count_weight, b.in_normal, b.in_albedo = ptr(in_color), ptr(in_weight), ptr(in_normal), ptr(in_albedo)
    b.out_color, b.out_sample_count_weight, b.out_normal, b.out_albedo = ptr(out_color), ptr(out_weight), ptr(out_normal), ptr(out_albedo)
    b.out_diagnostics = ptr(diagnostics)
    return b


def image_structs(images):
    """(ctypes array of rtb_image, the contiguous pixel arrays it points into)."""
    keep = []
    for im in images:
        a = np.ascontiguousarray(im, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("an image is a uint8 array [H, W, 3 or 4]")
        keep.append(a)
    arr = (abi.Image * max(len(keep), 1))()
    for i, a in enumerate(keep):
        arr[i].pixels, arr[i].width, arr[i].height, arr[i].pixel_stride = a.ctypes.data, a.shape[1], a.shape[0], a.shape[2]
    return arr, keep


def describe_scene(scene, leaf_spheres=1):
    """rtb_describe_scene: how the device would lay `scene` out (host-side only, no GPU needed)."""
    spheres = np.ascontiguousarray(scene.spheres, dtype=abi.SPHERE_DTYPE)
    materials = np.ascontiguousarray(scene.materials, dtype=abi.MATERIAL_DTYPE)
    nodes = np.ascontiguousarray(scene.nodes, dtype=abi.BVH_NODE_DTYPE)
    out = abi.SceneLayout()
    L = lib()
    rc = L.rtb_describe_scene(spheres.ctypes.data if len(spheres) else None, len(spheres),
                              materials.ctypes.data if len(materials) else None, len(materials),
                              nodes.ctypes.data if len(nodes) else None, len(nodes), int(leaf_spheres), C.byref(out))
    if rc != 0:
        raise RtbError(rc, (L.rtb_last_error(None) or b"").decode())
    return {name: getattr(out, name) for name, _ in abi.SceneLayout._fields_}


def retree_bvh(nodes):
    """rtb_retree_bvh: the topology the device walks under RTB_OPT_RETREE (host-side only, no GPU needed).
    -> BVH_NODE_DTYPE array (root = 0), or None when the world does not qualify."""
    nodes = np.ascontiguousarray(nodes, dtype=abi.BVH_NODE_DTYPE)
    out = np.zeros(max(len(nodes), 1), dtype=abi.BVH_NODE_DTYPE)
    count = C.c_size_t(0)
    L = lib()
    rc = L.rtb_retree_bvh(nodes.ctypes.data if len(nodes) else None, len(nodes), out.ctypes.data, len(out), C.byref(count))
    if rc != 0:
        raise RtbError(rc, (L.rtb_last_error(None) or b"").decode())
    return out[: count.value].copy() if count.value else None


class Context:
    """One rtb_ctx: a CUDA device, a stream, the uploaded world."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._L = lib()
        rc = self._L.rtb_create(device, C.byref(self._h))
        if rc != 0:
            raise RtbError(rc, (self._L.rtb_last_error(None) or b"").decode())
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise RtbError(rc, (self._L.rtb_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._L.rtb_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene -------------------------------------------------------------------------
    def upload_scene(self, spheres, materials, nodes):
        spheres = np.ascontiguousarray(spheres, dtype=abi.SPHERE_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=abi.MATERIAL_DTYPE)
        nodes = np.ascontiguousarray(nodes, dtype=abi.BVH_NODE_DTYPE)
        self._check(self._L.rtb_upload_scene(
            self._h, spheres.ctypes.data if len(spheres) else None, len(spheres),
            materials.ctypes.data if len(materials) else None, len(materials),
            nodes.ctypes.data if len(nodes) else None, len(nodes)))

    def upload_world(self, entities, spheres, triangles, materials, nodes):
        """rtb_upload_world: leaves of `nodes` index `entities`, each naming a sphere or a triangle."""
        entities = np.ascontiguousarray(entities, dtype=abi.ENTITY_DTYPE)
        spheres = np.ascontiguousarray(spheres, dtype=abi.SPHERE_DTYPE)
        triangles = np.ascontiguousarray(triangles, dtype=abi.TRIANGLE_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=abi.MATERIAL_DTYPE)
        nodes = np.ascontiguousarray(nodes, dtype=abi.BVH_NODE_DTYPE)

        def ptr(a):
            return a.ctypes.data if len(a) else None
        self._check(self._L.rtb_upload_world(self._h, ptr(entities), len(entities), ptr(spheres), len(spheres), ptr(triangles),
                                             len(triangles), ptr(materials), len(materials), ptr(nodes), len(nodes)))

    def upload_placed_world(self, entities, spheres, triangles, placed, materials, nodes):
        """rtb_upload_placed_world: upload_world plus entities with the reference's full Entity record (rotation,
        motion, Rect / Box content)."""
        entities = np.ascontiguousarray(entities, dtype=abi.ENTITY_DTYPE)
        spheres = np.ascontiguousarray(spheres, dtype=abi.SPHERE_DTYPE)
        triangles = np.ascontiguousarray(triangles, dtype=abi.TRIANGLE_DTYPE)
        placed = np.ascontiguousarray(placed, dtype=abi.PLACED_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=abi.MATERIAL_DTYPE)
        nodes = np.ascontiguousarray(nodes, dtype=abi.BVH_NODE_DTYPE)

        def ptr(a):
            return a.ctypes.data if len(a) else None
        self._check(self._L.rtb_upload_placed_world(self._h, ptr(entities), len(entities), ptr(spheres), len(spheres), ptr(triangles),
                                                    len(triangles), ptr(placed), len(placed), ptr(materials), len(materials),
                                                    ptr(nodes), len(nodes)))

    def upload_textures(self, images, material_textures, triangle_uvs=None):
        """rtb_upload_textures: `images` = list of uint8 arrays [H, W, 3 or 4]; `material_textures` = MATERIAL_TEXTURES_DTYPE per
        material of the uploaded world; `triangle_uvs` = float32 [n_triangles, 3, 2] or None."""
        imgs, keep = image_structs(images)
        mt = np.ascontiguousarray(material_textures, dtype=abi.MATERIAL_TEXTURES_DTYPE)
        uv = None if triangle_uvs is None else np.ascontiguousarray(triangle_uvs, dtype=np.float32)
        self._check(self._L.rtb_upload_textures(self._h, C.addressof(imgs) if len(keep) else None, len(keep), mt.ctypes.data if len(mt) else None,
                                                len(mt), uv.ctypes.data if uv is not None and uv.size else None,
                                                0 if uv is None else uv.size // 6))

    def upload_sky_cubemap(self, faces):
        """Environment.SkyCubemap: `faces` is a [6, H, W, 4] array of float16 (or their uint16 bits), +X -X +Y -Y +Z -Z;
        None removes it."""
        if faces is None:
            self._check(self._L.rtb_upload_sky_cubemap(self._h, None, 0, 0))
            return
        f = np.ascontiguousarray(faces)
        if f.dtype == np.float16:
            f = f.view(np.uint16)
        if f.dtype != np.uint16 or f.ndim != 4 or f.shape[0] != 6 or f.shape[3] != 4:
            raise ValueError("cubemap faces must be [6, H, W, 4] float16")
        self._check(self._L.rtb_upload_sky_cubemap(self._h, f.ctypes.data, f.shape[2], f.shape[1]))

    def upload(self, scene):
        if getattr(scene, "placed", None) is not None and len(scene.placed):
            self.upload_placed_world(scene.entities, scene.spheres, scene.triangles, scene.placed, scene.materials, scene.nodes)
        elif getattr(scene, "entities", None) is not None:
            self.upload_world(scene.entities, scene.spheres, scene.triangles, scene.materials, scene.nodes)
        else:
            self.upload_scene(scene.spheres, scene.materials, scene.nodes)
        if getattr(scene, "material_textures", None) is not None:
            self.upload_textures(scene.images, scene.material_textures, scene.triangle_uvs)

    # ---- the hot path ------------------------------------------------------------------
    def sample_batch(self, params, buffers, cancel=None):
        """Blocking call with HOST buffers (`HostBuffers`): H2D, megakernel, D2H."""
        b = buffers.as_struct()
        cancel_ptr = cancel.ctypes.data if cancel is not None else None
        self._check(self._L.rtb_sample_batch(self._h, C.byref(params), C.byref(b), cancel_ptr))
        return buffers

    def sample_batch_device(self, params, device_buffers, stream=None):
        """Enqueues the batch on device-resident buffers (`rtb_batch_buffers` of device pointers)."""
        self._check(self._L.rtb_sample_batch_device(self._h, C.byref(params), C.byref(device_buffers), stream))

    def register_host_buffers(self, buffers):
        for a in buffers.arrays():
            self._check(self._L.rtb_register_host_buffer(self._h, a.ctypes.data, a.nbytes))

    def unregister_host_buffers(self, buffers):
        for a in buffers.arrays():
            self._check(self._L.rtb_unregister_host_buffer(self._h, a.ctypes.data))

    # ---- adjacent jobs -----------------------------------------------------------------
    def combine_device(self, width, height, color4, normal3, albedo3, out_color3, out_normal3, out_albedo3,
                       debug_mode=False, ldr_albedo=False, stream=None):
        def ptr(x):
            return None if x is None else (x.data_ptr() if hasattr(x, "data_ptr") else int(x))
        self._check(self._L.rtb_combine_device(self._h, width, height, int(debug_mode), int(ldr_albedo), ptr(color4),
                                               ptr(normal3), ptr(albedo3), ptr(out_color3), ptr(out_normal3),
                                               ptr(out_albedo3), stream))

    def finalize_device(self, width, height, color3, normal3, albedo3, out_color, out_normal, out_albedo, stream=None):
        """FinalizeTexturesJob on device buffers: float3 images -> RGBA32 (one uint32 per pixel)."""
        def ptr(x):
            return None if x is None else (x.data_ptr() if hasattr(x, "data_ptr") else int(x))
        self._check(self._L.rtb_finalize_device(self._h, width, height, ptr(color3), ptr(normal3), ptr(albedo3), ptr(out_color),
                                                ptr(out_normal), ptr(out_albedo), stream))

    def reduce_metrics_device(self, width, height, diagnostics, color4, weight, stream=None):
        def ptr(x):
            return None if x is None else (x.data_ptr() if hasattr(x, "data_ptr") else int(x))
        m = abi.Metrics()
        self._check(self._L.rtb_reduce_metrics_device(self._h, width, height, ptr(diagnostics), ptr(color4), ptr(weight),
                                                      C.byref(m), stream))
        return m

    # ---- measurement -------------------------------------------------------------------
    def set_option(self, option, value):
        self._check(self._L.rtb_set_option(self._h, option, value))

    def counters(self):
        c = abi.Counters()
        self._check(self._L.rtb_get_counters(self._h, C.byref(c)))
        return {name: getattr(c, name) for name, _ in abi.Counters._fields_}

    def measure_fp32_peak(self, repeats=5):
        """Measured FP32 FMA peak of this device in TFLOP/s (the roofline that bounds this path)."""
        tf = C.c_double(0)
        self._check(self._L.rtb_measure_fp32_peak(self._h, repeats, C.byref(tf)))
        return tf.value

    def last_batch_in_place(self):
        """True when the last sample_batch ran on the (pinned) host arrays in place."""
        v = C.c_int(0)
        self._check(self._L.rtb_last_batch_in_place(self._h, C.byref(v)))
        return bool(v.value)

    def last_kernel_ms(self):
        ms = C.c_float(0)
        self._check(self._L.rtb_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


    # ---- device memory other rank processes can map (one process per GPU) -----------------------
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._L.rtb_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._check(self._L.rtb_device_free(self._h, ptr))

    def ipc_export(self, ptr):
        """-> 64 bytes another process of this box hands to `ipc_open`."""
        h = (C.c_ubyte * 64)()
        self._check(self._L.rtb_ipc_export(self._h, ptr, h))
        return bytes(h)

    def ipc_open(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self._L.rtb_ipc_open(self._h, buf, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._check(self._L.rtb_ipc_close(self._h, ptr))


def balance_rows(row_cost, row_begin, row_end, device_count):
    """rtb_balance_rows: the plugin's row-tile partition (host-side only, no GPU needed)."""
    cost = None if row_cost is None else np.ascontiguousarray(row_cost, dtype=np.float64)
    out = (C.c_int * (device_count + 1))()
    rc = lib().rtb_balance_rows(None if cost is None else cost.ctypes.data_as(C.POINTER(C.c_double)), row_begin, row_end, device_count, out)
    if rc != 0:
        raise RtbError(rc, (lib().rtb_last_error(None) or b"").decode())
    return list(out)


class MultiContext:
    """One rtb_multi: the frame of one sample job rendered by several GPUs of the box behind one call (no gather)."""

    def __init__(self, devices):
        self._L = lib()
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = self._L.rtb_multi_create(arr, len(devices), C.byref(self._h))
        if rc != 0:
            raise RtbError(rc, (self._L.rtb_multi_last_error(None) or b"").decode())
        self.devices = list(devices)

    def _check(self, rc):
        if rc != 0:
            raise RtbError(rc, (self._L.rtb_multi_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._L.rtb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option, value):
        self._check(self._L.rtb_multi_set_option(self._h, option, value))

    def upload(self, scene):
        def arr(a, dt):
            a = np.ascontiguousarray(a if a is not None else [], dtype=dt)
            return a, (a.ctypes.data if len(a) else None), len(a)
        sp, sp_p, sp_n = arr(scene.spheres, abi.SPHERE_DTYPE)
        ma, ma_p, ma_n = arr(scene.materials, abi.MATERIAL_DTYPE)
        no, no_p, no_n = arr(scene.nodes, abi.BVH_NODE_DTYPE)
        if getattr(scene, "entities", None) is not None:
            en, en_p, en_n = arr(scene.entities, abi.ENTITY_DTYPE)
            tr, tr_p, tr_n = arr(scene.triangles, abi.TRIANGLE_DTYPE)
            pl, pl_p, pl_n = arr(getattr(scene, "placed", None), abi.PLACED_DTYPE)
            self._check(self._L.rtb_multi_upload_placed_world(self._h, en_p, en_n, sp_p, sp_n, tr_p, tr_n, pl_p, pl_n, ma_p, ma_n, no_p, no_n))
        else:
            self._check(self._L.rtb_multi_upload_scene(self._h, sp_p, sp_n, ma_p, ma_n, no_p, no_n))
        if getattr(scene, "material_textures", None) is not None:
            imgs, keep = image_structs(scene.images)
            mt = np.ascontiguousarray(scene.material_textures, dtype=abi.MATERIAL_TEXTURES_DTYPE)
            uv = None if scene.triangle_uvs is None else np.ascontiguousarray(scene.triangle_uvs, dtype=np.float32)
            self._check(self._L.rtb_multi_upload_textures(self._h, C.addressof(imgs) if len(keep) else None, len(keep),
                                                          mt.ctypes.data if len(mt) else None, len(mt),
                                                          uv.ctypes.data if uv is not None and uv.size else None, 0 if uv is None else uv.size // 6))

    def register_host_buffers(self, buffers):
        for a in buffers.arrays():
            self._check(self._L.rtb_multi_register_host_buffer(self._h, a.ctypes.data, a.nbytes))

    def unregister_host_buffers(self, buffers):
        for a in buffers.arrays():
            self._check(self._L.rtb_multi_unregister_host_buffer(self._h, a.ctypes.data))

    def sample_batch(self, params, buffers, cancel=None):
        b = buffers.as_struct()
        self._check(self._L.rtb_multi_sample_batch(self._h, C.byref(params), C.byref(b), cancel.ctypes.data if cancel is not None else None))
        return buffers

    def sample_batch_device(self, params, device_buffers, owner_index=0, stream=None):
        self._check(self._L.rtb_multi_sample_batch_device(self._h, C.byref(params), C.byref(device_buffers), owner_index, stream))

    def tiles(self):
        """-> (row bounds [n + 1], kernel ms per device [n]) of the last batch."""
        n = len(self.devices)
        b = (C.c_int * (n + 1))()
        ms = (C.c_float * n)()
        self._check(self._L.rtb_multi_get_tiles(self._h, b, ms))
        return list(b), list(ms)
