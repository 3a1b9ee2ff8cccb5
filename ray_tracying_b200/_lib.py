"""ctypes binding of the C ABI in include/rt_render.h (librt_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ray_tracying_b200/csrc``.
There is no Python or CPU fallback: if the shared library is missing, importing this module
raises, and render calls fail with RT_ERR_CUDA when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB: another build of the same library (A/B runs of kernel variants, scripts/ab_variants.sh)
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")

RT_OK, RT_ERR_INVALID, RT_ERR_IO, RT_ERR_CUDA, RT_ERR_SCENE = 0, -1, -2, -3, -4
RT_SPHERE, RT_CUBE, RT_RECTANGLE, RT_PLANE = 0, 1, 2, 3


class CameraDesc(C.Structure):
    _fields_ = [
        ("location", C.c_float * 3), ("gaze", C.c_float * 3), ("up", C.c_float * 3),
        ("focal_length", C.c_float), ("sensor_width", C.c_int32), ("sensor_height", C.c_int32),
        ("aperture", C.c_float), ("focus_dist", C.c_float), ("res_x", C.c_int32), ("res_y", C.c_int32),
    ]


class LightDesc(C.Structure):
    _fields_ = [("location", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float), ("radius", C.c_float)]


class MaterialDesc(C.Structure):
    _fields_ = [
        ("diffuse_color", C.c_float * 3), ("specular_color", C.c_float * 3),
        ("k_ambient", C.c_float), ("k_diffuse", C.c_float), ("k_specular", C.c_float),
        ("shininess", C.c_float), ("roughness", C.c_float),
        ("reflectivity", C.c_float), ("transparency", C.c_float), ("refractive_index", C.c_float),
        ("texture", C.c_int32),
    ]


class ShapeDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("material", C.c_int32),
        ("translation", C.c_float * 3), ("rotation", C.c_float * 3), ("scale", C.c_float * 3),
        ("velocity", C.c_float * 3), ("corners", C.c_float * 12),
    ]


class TextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class SceneDesc(C.Structure):
    _fields_ = [
        ("camera", CameraDesc),
        ("n_lights", C.c_int32), ("lights", C.POINTER(LightDesc)),
        ("n_materials", C.c_int32), ("materials", C.POINTER(MaterialDesc)),
        ("n_shapes", C.c_int32), ("shapes", C.POINTER(ShapeDesc)),
        ("n_textures", C.c_int32), ("textures", C.POINTER(TextureDesc)),
    ]


class RenderParams(C.Structure):
    _fields_ = [
        ("use_bvh", C.c_int32), ("samples_sqrt", C.c_int32), ("light_samples", C.c_int32), ("max_depth", C.c_int32),
        ("seed", C.c_uint64), ("fixed_time", C.c_float),
        ("rank", C.c_int32), ("world", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32),
        ("collect_stats", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class RenderStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("secondary_rays", C.c_uint64),
        ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
        ("kernel_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_int32), ("pixels", C.c_int32),
    ]


class BvhNodeDump(C.Structure):
    _fields_ = [
        ("is_leaf", C.c_int32), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
        ("count", C.c_int32), ("prims", C.c_int32 * 4),
    ]


# every symbol include/rt_render.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "rt_last_error": (C.c_char_p, []),
    "rt_version": (C.c_int, []),
    "rt_device_count": (C.c_int, []),
    "rt_render_params_default": (None, [C.POINTER(RenderParams)]),
    "rt_scene_load_json": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_VP)]),
    "rt_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(_VP)]),
    "rt_scene_destroy": (None, [_VP]),
    "rt_scene_resolution": (C.c_int, [_VP, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rt_scene_counts": (C.c_int, [_VP] + [C.POINTER(C.c_int32)] * 5),
    "rt_scene_shape_order": (C.c_int, [_VP, C.POINTER(C.c_int32), C.c_int32]),
    "rt_scene_dump_bvh": (C.c_int, [_VP, C.POINTER(BvhNodeDump), C.c_int32]),
    "rt_scene_upload": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "rt_scene_evict": (C.c_int, [_VP]),
    "rt_scene_last_timing": (C.c_int, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rt_shard_pixels": (C.c_int, [_VP, C.POINTER(RenderParams), C.POINTER(C.c_int64)]),
    "rt_render_device": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _VP, _VP, _VP, C.POINTER(RenderStats)]),
    "rt_selftest_boxes": (C.c_int, [C.c_uint64, C.c_int64, C.POINTER(C.c_uint64)]),
    "rt_selftest_cull": (C.c_int, [_VP, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64)]),
    "rt_json_number": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "rt_scene_dump_wide": (C.c_int, [_VP, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]),
    "rt_scene_last_kernel_times": (C.c_int, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rt_render": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _VP, _VP, C.POINTER(RenderStats)]),
    "rt_render_multi": (C.c_int, [_VP, C.POINTER(RenderParams), C.c_int32, C.POINTER(C.c_int32), _VP, _VP, _VP, C.POINTER(RenderStats)]),
    "rt_scene_launch_count": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "rt_traversal_peak": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "rt_write_ppm": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, _VP]),
    "rt_read_ppm": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_uint8))]),
    "rt_free": (None, [_VP]),
}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the render path)"
        )
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = load_library()


class RtError(RuntimeError):
    def __init__(self, status: int, where: str):
        msg = lib.rt_last_error().decode("utf-8", "replace")
        super().__init__(f"{where} failed with status {status}: {msg}")
        self.status = status


def check(status: int, where: str) -> None:
    if status != RT_OK:
        raise RtError(status, where)
