"""ctypes / numpy mirrors of the PODs in include/rtb.h and include/rtb_host.h.

Field order and sizes must match the headers byte for byte; tests/test_abi.py checks
sizeof() of every struct against a table compiled from the headers.
"""
import ctypes as C

import numpy as np

# ---- numpy record dtypes for the scene arrays (rtb.h) ---------------------------------------
SPHERE_DTYPE = np.dtype(
    [("center", "<f4", 3), ("radius", "<f4"), ("material", "<u4"), ("reserved", "<u4", 3)], align=False
)  # rtb_sphere, 32 B
MATERIAL_DTYPE = np.dtype(
    [
        ("type", "<u4"),
        ("albedo", "<f4", 3),
        ("emission", "<f4", 3),
        ("glossiness", "<f4"),
        ("metallic", "<f4"),
        ("index_of_refraction", "<f4"),
        ("reserved", "<u4", 2),
    ],
    align=False,
)  # rtb_material, 48 B
BVH_NODE_DTYPE = np.dtype(
    [
        ("bounds_min", "<f4", 3),
        ("bounds_max", "<f4", 3),
        ("left", "<i4"),
        ("right", "<i4"),
        ("first_entity", "<i4"),
        ("entity_count", "<i4"),
    ],
    align=False,
)  # rtb_bvh_node, 40 B
DIAGNOSTICS_DTYPE = np.dtype(
    [("ray_count", "<f4"), ("bounds_hit_count", "<f4"), ("candidate_count", "<f4"), ("sample_count_weight", "<f4")]
)  # rtb_diagnostics, 16 B

TRIANGLE_DTYPE = np.dtype(
    [("edge2", "<f4", 3), ("edge1", "<f4", 3), ("v0", "<f4", 3), ("normals", "<f4", (3, 3)), ("material", "<u4"), ("reserved", "<u4")],
    align=False,
)  # rtb_triangle, 80 B
ENTITY_DTYPE = np.dtype([("type", "<u4"), ("index", "<u4")])  # rtb_entity, 8 B
ENTITY_SPHERE, ENTITY_RECT, ENTITY_BOX, ENTITY_TRIANGLE = 1, 2, 3, 4
ENTITY_PLACED = 0x100  # flag: rtb_entity.index points into the placed-entity array
PLACED_DTYPE = np.dtype(
    [("type", "<u4"), ("material", "<u4"), ("moving", "<u4"), ("reserved", "<u4"), ("rotation", "<f4", 4), ("position", "<f4", 3),
     ("destination_offset", "<f4", 3), ("time_range", "<f4", 2), ("size", "<f4", 3), ("reserved2", "<f4")],
    align=False,
)  # rtb_placed_entity, 80 B
assert PLACED_DTYPE.itemsize == 80


class Image(C.Structure):  # rtb_image
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("pixel_stride", C.c_int32), ("reserved", C.c_int32)]


MATERIAL_TEXTURES_DTYPE = np.dtype(
    [("albedo_image", "<i4"), ("emission_image", "<i4"), ("glossiness_image", "<i4"), ("metallic_image", "<i4"),
     ("glossiness_channel", "<i4"), ("metallic_channel", "<i4"), ("reserved", "<i4", 2)]
)  # rtb_material_textures, 32 B
assert MATERIAL_TEXTURES_DTYPE.itemsize == 32

assert SPHERE_DTYPE.itemsize == 32 and MATERIAL_DTYPE.itemsize == 48 and BVH_NODE_DTYPE.itemsize == 40
assert TRIANGLE_DTYPE.itemsize == 80 and ENTITY_DTYPE.itemsize == 8

MATERIAL_STANDARD, MATERIAL_DIELECTRIC, MATERIAL_PROBABILISTIC_VOLUME = 0, 1, 2
SKY_NONE, SKY_GRADIENT, SKY_CUBEMAP = 0, 1, 2
SCENE_THREE_SPHERES, SCENE_FINAL, SCENE_STRESS = 0, 1, 2

ABI_VERSION = 2
RTB_OK = 0
RTB_ERR_INVALID_ARGUMENT = 1
RTB_ERR_NO_SCENE = 2
RTB_ERR_CANCELLED = 3
RTB_ERR_UNSUPPORTED = 4
RTB_ERR_OUT_OF_MEMORY = 5
RTB_ERR_CUDA = 100

OPT_COUNTERS, OPT_KERNEL, OPT_LEAF_SPHERES, OPT_ALWAYS_WALK_CHAINS, OPT_HOST_ACCESS, OPT_NOISE, OPT_BALANCE_TILES, OPT_MATH, OPT_RETREE = 1, 2, 4, 5, 6, 7, 8, 9, 10
MATH_PARITY, MATH_FAST = 0, 1
NOISE_PHILOX, NOISE_WHITE = 0, 1
KERNEL_AUTO, KERNEL_SIMPLE, KERNEL_MEGA = 0, 1, 2

f32 = C.c_float
f32x2 = C.c_float * 2
f32x3 = C.c_float * 3


class View(C.Structure):  # rtb_view == Runtime/View.cs:8-14
    _fields_ = [
        ("origin", f32x3),
        ("lower_left_corner", f32x3),
        ("horizontal", f32x3),
        ("vertical", f32x3),
        ("forward", f32x3),
        ("up", f32x3),
        ("right", f32x3),
        ("lens_radius", f32),
    ]


class Environment(C.Structure):  # rtb_environment == Runtime/Environment.cs:12-18
    _fields_ = [("sky_type", C.c_uint32), ("sky_bottom_color", f32x3), ("sky_top_color", f32x3)]


class BatchParams(C.Structure):  # rtb_batch_params == the uniform fields of SampleBatchJob.cs:25-39
    _fields_ = [
        ("size", f32x2),
        ("slice_offset", C.c_int32),
        ("slice_divider", C.c_int32),
        ("seed", C.c_uint32),
        ("view", View),
        ("environment", Environment),
        ("sample_count_range", C.c_uint32 * 2),
        ("trace_depth", C.c_int32),
        ("sub_pixel_jitter", C.c_uint32),
        ("sample_count_weight_extrema", f32x2),
        ("row_begin", C.c_int32),
        ("row_end", C.c_int32),
    ]


class BatchBuffers(C.Structure):  # rtb_batch_buffers == the NativeArray fields of SampleBatchJob.cs:41-51
    _fields_ = [
        ("in_color", C.c_void_p),
        ("in_sample_count_weight", C.c_void_p),
        ("in_normal", C.c_void_p),
        ("in_albedo", C.c_void_p),
        ("out_color", C.c_void_p),
        ("out_sample_count_weight", C.c_void_p),
        ("out_normal", C.c_void_p),
        ("out_albedo", C.c_void_p),
        ("out_diagnostics", C.c_void_p),
    ]


class Metrics(C.Structure):  # rtb_metrics == ReduceMetricsJob.cs:17-20
    _fields_ = [
        ("total_ray_count", C.c_int64),
        ("total_samples", C.c_int64),
        ("sample_count_weight_min", f32),
        ("sample_count_weight_max", f32),
        ("sample_count_min", C.c_int32),
        ("sample_count_max", C.c_int32),
    ]


class Counters(C.Structure):  # rtb_counters
    _fields_ = [
        ("samples", C.c_uint64),
        ("rays", C.c_uint64),
        ("node_tests", C.c_uint64),
        ("sphere_tests", C.c_uint64),
        ("shade_standard", C.c_uint64),
        ("shade_dielectric", C.c_uint64),
        ("sky_hits", C.c_uint64),
        ("failed_samples", C.c_uint64),
    ]


class SceneLayout(C.Structure):  # rtb_scene_layout
    _fields_ = [(n, C.c_uint32) for n in ("inner_nodes", "leaves", "device_spheres", "max_leaf_spheres", "max_depth",
                                          "blob_bytes", "chain_boxes", "collapsed")]


class Camera(C.Structure):  # rtbh_camera
    _fields_ = [("position", f32x3), ("target", f32x3), ("aperture", f32), ("vertical_fov", f32)]


class SceneInfo(C.Structure):  # rtbh_scene_info
    _fields_ = [
        ("camera", Camera),
        ("environment", Environment),
        ("sphere_count", C.c_uint32),
        ("material_count", C.c_uint32),
        ("lambertian_count", C.c_uint32),
        ("metal_count", C.c_uint32),
        ("dielectric_count", C.c_uint32),
        ("tentative_draws", C.c_uint32),
    ]


class URandom(C.Structure):  # rtbh_random
    _fields_ = [("state", C.c_uint32)]


STRUCT_SIZES = {  # name in the headers -> python mirror; checked against sizeof() from C in tests
    "rtb_sphere": SPHERE_DTYPE.itemsize,
    "rtb_material": MATERIAL_DTYPE.itemsize,
    "rtb_bvh_node": BVH_NODE_DTYPE.itemsize,
    "rtb_triangle": TRIANGLE_DTYPE.itemsize,
    "rtb_entity": ENTITY_DTYPE.itemsize,
    "rtb_placed_entity": PLACED_DTYPE.itemsize,
    "rtb_image": C.sizeof(Image),
    "rtb_material_textures": MATERIAL_TEXTURES_DTYPE.itemsize,
    "rtb_diagnostics": DIAGNOSTICS_DTYPE.itemsize,
    "rtb_view": C.sizeof(View),
    "rtb_environment": C.sizeof(Environment),
    "rtb_batch_params": C.sizeof(BatchParams),
    "rtb_batch_buffers": C.sizeof(BatchBuffers),
    "rtb_metrics": C.sizeof(Metrics),
    "rtb_counters": C.sizeof(Counters),
    "rtb_scene_layout": C.sizeof(SceneLayout),
    "rtbh_camera": C.sizeof(Camera),
    "rtbh_scene_info": C.sizeof(SceneInfo),
}
