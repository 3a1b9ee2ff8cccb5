"""ray_tracying_b200 -- B200-native (sm_100a) implementation of the EricZhang12138/Ray_Tracying
hot path: per-pixel ray generation, BVH traversal, ray-primitive intersection and Blinn-Phong
shading with shadow, reflection and refraction rays, behind a C ABI (include/rt_render.h).

Importing this package loads ``librt_b200.so``; it raises if the library has not been built.
"""
from .api import (MAX_RECURSION_DEPTH, MATERIAL_DTYPE, SHAPE_DTYPE, Scene, Stats, device_count, make_params, read_ppm,
                  selftest_boxes, write_ppm)
from ._lib import RT_CUBE, RT_PLANE, RT_RECTANGLE, RT_SPHERE, RtError

__all__ = [
    "MAX_RECURSION_DEPTH", "MATERIAL_DTYPE", "SHAPE_DTYPE", "Scene", "Stats", "device_count", "make_params",
    "read_ppm", "write_ppm", "selftest_boxes", "RT_SPHERE", "RT_CUBE", "RT_RECTANGLE", "RT_PLANE", "RtError",
]
