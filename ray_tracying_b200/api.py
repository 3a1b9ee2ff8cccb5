"""Host-side mirror of the reference's front end, over the C ABI.

``Scene.from_json`` replaces Camera() + load_lights_from_json() + load_shapes_from_json() + BVH()
(reference Code/raytracer.cpp:400-422); ``Scene.render`` replaces the frame loop
(raytracer.cpp:433-476) with the same switches as the command line (-bvh, -s, -light_sample).
Argument names follow the reference's CLI variables (use_bvh, n_samples_sqrt, light_samples).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (BvhNodeDump, CameraDesc, LightDesc, MaterialDesc, RenderParams, RenderStats, SceneDesc, ShapeDesc,
                   TextureDesc, check, lib)

MAX_RECURSION_DEPTH = 10  # reference Code/raytracer.hpp:11


@dataclass
class Stats:
    rays: int
    primary_rays: int
    shadow_rays: int
    secondary_rays: int
    node_visits: int
    prim_tests: int
    kernel_ms: float
    total_ms: float
    launches: int
    pixels: int

    @staticmethod
    def from_c(s: RenderStats) -> "Stats":
        return Stats(s.rays, s.primary_rays, s.shadow_rays, s.secondary_rays, s.node_visits, s.prim_tests,
                     s.kernel_ms, s.total_ms, s.launches, s.pixels)


def make_params(use_bvh: bool = False, n_samples_sqrt: int = 4, light_samples: int = 1,
                max_depth: int = MAX_RECURSION_DEPTH, seed: int = 1, fixed_time: float = -1.0, rank: int = 0,
                world: int = 1, tile: Sequence[int] = (32, 32), collect_stats: bool = False,
                prune: bool = True, time_kernels: bool = False, serial: bool = False,
                window: Optional[Sequence[int]] = None, tile_block: int = 0) -> RenderParams:
    """Defaults are the reference binary's (raytracer.cpp:361-363: BVH off, 4x4 samples, 1 light sample).
    window = (x0, y0, x1, y1): render only that region of the frame; tile_block = B: deal tiles to the
    ranks in B x B groups."""
    p = RenderParams()
    lib.rt_render_params_default(C.byref(p))
    p.use_bvh = int(bool(use_bvh))
    p.samples_sqrt = int(n_samples_sqrt)
    p.light_samples = int(light_samples)
    p.max_depth = int(max_depth)
    p.seed = int(seed)
    p.fixed_time = float(fixed_time)
    p.rank, p.world = int(rank), int(world)
    p.tile_w, p.tile_h = int(tile[0]), int(tile[1])
    p.collect_stats = int(bool(collect_stats))
    p.reserved[0] = 0 if prune else 1
    p.reserved[1] = 1 if time_kernels else 0
    p.reserved[2] = 1 if serial else 0
    if window is not None:
        x0, y0, x1, y1 = (int(v) for v in window)
        if not (0 <= x0 < x1 <= 65535 and 0 <= y0 < y1 <= 65535):
            raise ValueError("window must satisfy 0 <= x0 < x1 <= 65535 and 0 <= y0 < y1 <= 65535")
        p.reserved[3] = C.c_int32(x0 | (x1 << 16)).value
        p.reserved[4] = C.c_int32(y0 | (y1 << 16)).value
    p.reserved[5] = int(tile_block)
    return p


class Scene:
    """A loaded scene: host model + the reference's BVH, flattened; device copy made on demand."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def from_json(cls, scene_path: str, texture_dir: Optional[str] = None) -> "Scene":
        out = C.c_void_p()
        check(lib.rt_scene_load_json(scene_path.encode(), texture_dir.encode() if texture_dir else None, C.byref(out)),
              "rt_scene_load_json")
        return cls(out.value)

    @classmethod
    def from_arrays(cls, camera: dict, lights: np.ndarray, materials: np.ndarray, shapes: np.ndarray,
                    textures: Sequence[np.ndarray] = ()) -> "Scene":
        """Constructor-level creation.

        camera   : dict with the rt_camera_desc fields
        lights   : (L, 8) float32  = location[3], color[3], intensity, radius
        materials: (M, 15) float32 = diffuse[3], specular[3], ka, kd, ks, shininess, roughness,
                   reflectivity, transparency, refractive_index, texture index (-1 = none)
        shapes   : structured array with dtype SHAPE_DTYPE
        textures : sequence of (H, W, 3) uint8 arrays
        """
        desc = SceneDesc()
        cam = desc.camera
        for k in ("location", "gaze", "up"):
            for i in range(3):
                getattr(cam, k)[i] = float(camera[k][i])
        cam.focal_length = float(camera["focal_length"])
        cam.sensor_width = int(camera["sensor_width"])
        cam.sensor_height = int(camera["sensor_height"])
        cam.aperture = float(camera.get("aperture", 0.0))
        cam.focus_dist = float(camera.get("focus_dist", 10.0))
        cam.res_x, cam.res_y = int(camera["res_x"]), int(camera["res_y"])

        lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1, 8)
        materials = np.ascontiguousarray(materials, dtype=np.float32).reshape(-1, 15)
        mats_c = np.zeros(len(materials), dtype=MATERIAL_DTYPE)
        mats_c["f"] = materials[:, :14]
        mats_c["texture"] = materials[:, 14].astype(np.int32)
        shapes = np.ascontiguousarray(shapes, dtype=SHAPE_DTYPE)
        assert LightDesc and C.sizeof(LightDesc) == 32 and C.sizeof(MaterialDesc) == MATERIAL_DTYPE.itemsize
        assert C.sizeof(ShapeDesc) == SHAPE_DTYPE.itemsize

        tex_arrays = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        tex_c = (TextureDesc * max(1, len(tex_arrays)))()
        for i, t in enumerate(tex_arrays):
            tex_c[i].height, tex_c[i].width = t.shape[0], t.shape[1]
            tex_c[i].rgb = t.ctypes.data_as(C.POINTER(C.c_uint8))

        desc.n_lights = len(lights)
        desc.lights = lights.ctypes.data_as(C.POINTER(LightDesc))
        desc.n_materials = len(mats_c)
        desc.materials = mats_c.ctypes.data_as(C.POINTER(MaterialDesc))
        desc.n_shapes = len(shapes)
        desc.shapes = shapes.ctypes.data_as(C.POINTER(ShapeDesc))
        desc.n_textures = len(tex_arrays)
        desc.textures = tex_c
        out = C.c_void_p()
        check(lib.rt_scene_create(C.byref(desc), C.byref(out)), "rt_scene_create")
        return cls(out.value)

    def close(self) -> None:
        if self._h:
            lib.rt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ------------------------------------------------------------------------
    @property
    def resolution(self) -> tuple[int, int]:
        w, h = C.c_int32(), C.c_int32()
        check(lib.rt_scene_resolution(self._h, C.byref(w), C.byref(h)), "rt_scene_resolution")
        return w.value, h.value

    def counts(self) -> dict:
        v = [C.c_int32() for _ in range(5)]
        check(lib.rt_scene_counts(self._h, *[C.byref(x) for x in v]), "rt_scene_counts")
        return dict(zip(("shapes", "lights", "materials", "nodes", "leaves"), (x.value for x in v)))

    def shape_order(self) -> np.ndarray:
        n = self.counts()["shapes"]
        out = np.zeros(n, dtype=np.int32)
        check(lib.rt_scene_shape_order(self._h, out.ctypes.data_as(C.POINTER(C.c_int32)), n), "rt_scene_shape_order")
        return out

    def dump_bvh(self) -> list:
        n = self.counts()["nodes"]
        buf = (BvhNodeDump * max(1, n))()
        got = lib.rt_scene_dump_bvh(self._h, buf, n)
        if got < 0:
            check(got, "rt_scene_dump_bvh")
        return [(b.is_leaf, tuple(b.box_min), tuple(b.box_max), [b.prims[k] for k in range(b.count)]) for b in buf[:got]]

    def selftest_cull(self, rays_per_primitive: int = 64, seed: int = 1) -> dict:
        """rt_selftest_cull: exact primitive hits vs the conservative test of the culling boxes."""
        out = (C.c_uint64 * 8)()
        check(lib.rt_selftest_cull(self._h, seed, rays_per_primitive, out), "rt_selftest_cull")
        return dict(zip(("tests", "hits", "passes", "violations"), (int(v) for v in out)))

    def dump_wide(self):
        """(nodes, depth): the 4-wide device tree as an (N, 32) float32 array (csrc/scene.hpp DWide) and its depth."""
        depth = C.c_int32()
        n = lib.rt_scene_dump_wide(self._h, None, 0, C.byref(depth))
        if n < 0:
            check(n, "rt_scene_dump_wide")
        out = np.zeros((max(n, 1), 32), dtype=np.float32)
        check(min(0, lib.rt_scene_dump_wide(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(depth))), "rt_scene_dump_wide")
        return out[:n], depth.value

    # -- device -------------------------------------------------------------------------------
    def upload(self) -> int:
        n = C.c_uint64()
        check(lib.rt_scene_upload(self._h, C.byref(n)), "rt_scene_upload")
        return n.value

    def evict(self) -> None:
        check(lib.rt_scene_evict(self._h), "rt_scene_evict")

    def last_timing(self) -> tuple[float, float]:
        """(kernel_ms, total_ms) of the most recent render call, from CUDA events on its stream."""
        k, t = C.c_float(), C.c_float()
        check(lib.rt_scene_last_timing(self._h, C.byref(k), C.byref(t)), "rt_scene_last_timing")
        return k.value, t.value

    def last_kernel_times(self) -> dict:
        """{class: (ms, timed launches, launches of the frame)} of the most recent frame rendered with
        time_kernels=True (at most 512 launches of a frame are timed)."""
        ms, n, tot = (C.c_float * 4)(), (C.c_int32 * 4)(), (C.c_int32 * 4)()
        check(lib.rt_scene_last_kernel_times(self._h, ms, n, tot), "rt_scene_last_kernel_times")
        return {k: (ms[i], n[i], tot[i]) for i, k in enumerate(("trace", "shadow", "shade", "light"))}

    def shard_pixels(self, params: RenderParams) -> int:
        n = C.c_int64()
        check(lib.rt_shard_pixels(self._h, C.byref(params), C.byref(n)), "rt_shard_pixels")
        return n.value

    def render(self, params: Optional[RenderParams] = None, want_ids: bool = True, want_linear: bool = False, **kw):
        """Host-buffer render (rt_render): returns (rgb8 HxWx3 uint8, ids HxW int32 | None, linear | None, Stats)."""
        p = params if params is not None else make_params(**kw)
        w, h = self.resolution
        rgb = np.zeros((h, w, 3), dtype=np.uint8)
        ids = np.full((h, w), -1, dtype=np.int32) if want_ids else None
        lin = np.zeros((h, w, 3), dtype=np.float32) if want_linear else None
        st = RenderStats()
        check(lib.rt_render(self._h, C.byref(p), rgb.ctypes.data, ids.ctypes.data if want_ids else None,
                            lin.ctypes.data if want_linear else None, C.byref(st)), "rt_render")
        return rgb, ids, lin, Stats.from_c(st)

    def render_multi(self, params: Optional[RenderParams] = None, n_devices: int = 1, devices: Optional[Sequence[int]] = None,
                     want_ids: bool = True, want_linear: bool = False, **kw):
        """rt_render_multi: the whole frame over several GPUs of this process (tiles dealt to the devices, scene
        replicated, each device's tiles copied to the host at frame end). Same return value as render()."""
        p = params if params is not None else make_params(**kw)
        w, h = self.resolution
        rgb = np.zeros((h, w, 3), dtype=np.uint8)
        ids = np.full((h, w), -1, dtype=np.int32) if want_ids else None
        lin = np.zeros((h, w, 3), dtype=np.float32) if want_linear else None
        dev = (C.c_int32 * n_devices)(*devices) if devices is not None else None
        st = RenderStats()
        check(lib.rt_render_multi(self._h, C.byref(p), n_devices, dev, rgb.ctypes.data, ids.ctypes.data if want_ids else None,
                                  lin.ctypes.data if want_linear else None, C.byref(st)), "rt_render_multi")
        return rgb, ids, lin, Stats.from_c(st)

    def render_multi_into(self, params: RenderParams, n_devices: int, rgb_ptr: int = 0, ids_ptr: int = 0, linear_ptr: int = 0) -> Stats:
        """rt_render_multi into caller-owned HOST memory given as raw addresses (page-locked or not)."""
        st = RenderStats()
        check(lib.rt_render_multi(self._h, C.byref(params), n_devices, None, rgb_ptr or None, ids_ptr or None, linear_ptr or None,
                                  C.byref(st)), "rt_render_multi")
        return Stats.from_c(st)

    def launch_count(self) -> int:
        """Kernels launched for this scene so far, over all devices (rt_scene_launch_count)."""
        n = C.c_uint64()
        check(lib.rt_scene_launch_count(self._h, C.byref(n)), "rt_scene_launch_count")
        return n.value

    def traversal_peak(self, any_hit: bool = False, n_nodes: int = 64, steps: int = 4096, repeats: int = 5) -> tuple[float, float]:
        """(box tests per second, ms of the best launch) of the traversal loop's node step with fully converged
        warps on L1-resident nodes of this scene's tree (rt_traversal_peak): the traversal roofline."""
        v, ms = C.c_double(), C.c_float()
        check(lib.rt_traversal_peak(self._h, int(any_hit), n_nodes, steps, repeats, C.byref(v), C.byref(ms)), "rt_traversal_peak")
        return v.value, ms.value

    def render_into(self, params: RenderParams, rgb_ptr: int = 0, ids_ptr: int = 0, linear_ptr: int = 0) -> Stats:
        """Host-buffer render (rt_render) into caller-owned HOST memory given as raw addresses (e.g. a
        page-locked torch tensor's .data_ptr()): upload if the device copy is stale, render, copy back."""
        st = RenderStats()
        check(lib.rt_render(self._h, C.byref(params), rgb_ptr or None, ids_ptr or None, linear_ptr or None, C.byref(st)),
              "rt_render")
        return Stats.from_c(st)

    def render_device(self, params: RenderParams, rgb_ptr: int = 0, ids_ptr: int = 0, linear_ptr: int = 0,
                      stream: int = 0, sync_stats: bool = True) -> Optional[Stats]:
        """Device-buffer render (rt_render_device) on raw device pointers (e.g. torch .data_ptr())."""
        st = RenderStats()
        check(lib.rt_render_device(self._h, C.byref(p := params), rgb_ptr or None, ids_ptr or None, linear_ptr or None,
                                   stream or None, C.byref(st) if sync_stats else None), "rt_render_device")
        del p
        return Stats.from_c(st) if sync_stats else None


SHAPE_DTYPE = np.dtype([
    ("type", np.int32), ("material", np.int32),
    ("translation", np.float32, 3), ("rotation", np.float32, 3), ("scale", np.float32, 3),
    ("velocity", np.float32, 3), ("corners", np.float32, 12),
])
MATERIAL_DTYPE = np.dtype([("f", np.float32, 14), ("texture", np.int32)])


def write_ppm(path: str, rgb: np.ndarray) -> None:
    """Image::write (reference Code/image.cpp:53-84): ASCII P3."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    check(lib.rt_write_ppm(path.encode(), rgb.shape[1], rgb.shape[0], rgb.ctypes.data), "rt_write_ppm")


def read_ppm(path: str) -> np.ndarray:
    """Image::read (reference Code/image.cpp:86-133)."""
    w, h = C.c_int32(), C.c_int32()
    ptr = C.POINTER(C.c_uint8)()
    check(lib.rt_read_ppm(path.encode(), C.byref(w), C.byref(h), C.byref(ptr)), "rt_read_ppm")
    try:
        return np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.rt_free(ptr)


def selftest_boxes(n: int, seed: int = 1) -> dict:
    """rt_selftest_boxes: conservative slab test vs the reference's exact box test on n random pairs."""
    out = (C.c_uint64 * 8)()
    check(lib.rt_selftest_boxes(seed, n, out), "rt_selftest_boxes")
    keys = ("tests", "exact", "conservative", "surely", "violations_exact_not_conservative", "violations_surely_not_exact", "skipped")
    return dict(zip(keys, (int(v) for v in out)))


def device_count() -> int:
    return lib.rt_device_count()
