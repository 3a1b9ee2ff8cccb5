"""ray_tracying_b200 -- B200-native (sm_100a) implementation of the EricZhang12138/Ray_Tracying
hot path: per-pixel ray generation, BVH traversal, ray-primitive intersection and Blinn-Phong
shading with shadow, reflection and refraction rays, behind a C ABI (include/rt_render.h).

``librt_b200.so`` is loaded the first time anything of the render API is touched (``rt.Scene``,
``rt.make_params``, ...); that raises if the library has not been built -- there is no Python or
CPU fallback. The pure-Python helpers (``scenes``, ``workloads``, ``dist``) import without it, so
that the CPU reference arm of bench.py never maps the product library.
"""
_API = ("MAX_RECURSION_DEPTH", "MATERIAL_DTYPE", "SHAPE_DTYPE", "Scene", "Stats", "device_count", "make_params",
        "read_ppm", "write_ppm", "selftest_boxes")
_LIB = ("RT_SPHERE", "RT_CUBE", "RT_RECTANGLE", "RT_PLANE", "RtError")

__all__ = list(_API + _LIB)


def __getattr__(name):
    if name in _API:
        from . import api
        return getattr(api, name)
    if name in _LIB:
        from . import _lib
        return getattr(_lib, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def require_library() -> None:
    """Loads librt_b200.so now (ImportError if it is missing or lacks a declared symbol)."""
    from . import _lib  # noqa: F401
