"""oracle/oracle.py -- TEST INFRASTRUCTURE: Python access to the two CPU checkers.

* ``OracleScene``  : ctypes binding of oracle/librt_oracle.so (our CPU restatement, rt_oracle.cpp)
* ``RefDriver``    : runs oracle/_ref/ref_driver, the UNMODIFIED reference compiled from
                     /root/reference (present only where it was built; see oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

from . import scene_io

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "librt_oracle.so")
REF_DRIVER = os.path.join(_HERE, "_ref", "ref_driver")


class _Camera(C.Structure):
    _fields_ = [("location", C.c_float * 3), ("gaze", C.c_float * 3), ("up", C.c_float * 3), ("focal_length", C.c_float),
                ("sensor_width", C.c_int32), ("sensor_height", C.c_int32), ("aperture", C.c_float),
                ("focus_dist", C.c_float), ("res_x", C.c_int32), ("res_y", C.c_int32)]


class _Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class _SceneDesc(C.Structure):
    _fields_ = [("camera", _Camera), ("n_lights", C.c_int32), ("lights", C.c_void_p), ("n_materials", C.c_int32),
                ("materials", C.c_void_p), ("n_shapes", C.c_int32), ("shapes", C.c_void_p), ("n_textures", C.c_int32),
                ("textures", C.POINTER(_Texture))]


class _Params(C.Structure):
    _fields_ = [("use_bvh", C.c_int32), ("samples_sqrt", C.c_int32), ("light_samples", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint64), ("fixed_time", C.c_float), ("row0", C.c_int32), ("row1", C.c_int32),
                ("threads", C.c_int32), ("col0", C.c_int32), ("col1", C.c_int32)]


class _NodeDump(C.Structure):
    _fields_ = [("is_leaf", C.c_int32), ("lo", C.c_float * 3), ("hi", C.c_float * 3), ("count", C.c_int32),
                ("prims", C.c_int32 * 4)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            raise RuntimeError(f"{ORACLE_LIB} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(ORACLE_LIB)
        _lib.orc_scene_create.argtypes = [C.POINTER(_SceneDesc), C.POINTER(C.c_void_p)]
        _lib.orc_scene_destroy.argtypes = [C.c_void_p]
        _lib.orc_scene_shape_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        _lib.orc_scene_dump_bvh.argtypes = [C.c_void_p, C.POINTER(_NodeDump), C.c_int32]
        _lib.orc_render.argtypes = [C.c_void_p, C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def philox4x32_10(ctr, key):
    lib = _load()
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


class OracleScene:
    def __init__(self, camera: dict, lights, materials, shapes, textures=()):
        lib = _load()
        self._keep = (np.ascontiguousarray(lights), np.ascontiguousarray(materials), np.ascontiguousarray(shapes),
                      [np.ascontiguousarray(t, dtype=np.uint8) for t in textures])
        lights, materials, shapes, textures = self._keep
        assert lights.dtype.itemsize == 32 and materials.dtype.itemsize == 60 and shapes.dtype.itemsize == 104
        d = _SceneDesc()
        for k in ("location", "gaze", "up"):
            for i in range(3):
                getattr(d.camera, k)[i] = float(camera[k][i])
        d.camera.focal_length = float(camera["focal_length"])
        d.camera.sensor_width, d.camera.sensor_height = int(camera["sensor_width"]), int(camera["sensor_height"])
        d.camera.aperture, d.camera.focus_dist = float(camera["aperture"]), float(camera["focus_dist"])
        d.camera.res_x, d.camera.res_y = int(camera["res_x"]), int(camera["res_y"])
        d.n_lights, d.lights = len(lights), lights.ctypes.data
        d.n_materials, d.materials = len(materials), materials.ctypes.data
        d.n_shapes, d.shapes = len(shapes), shapes.ctypes.data
        tex = (_Texture * max(1, len(textures)))()
        for i, t in enumerate(textures):
            tex[i].height, tex[i].width = t.shape[0], t.shape[1]
            tex[i].rgb = t.ctypes.data_as(C.POINTER(C.c_uint8))
        d.n_textures, d.textures = len(textures), tex
        self.width, self.height = d.camera.res_x, d.camera.res_y
        self.n_shapes = len(shapes)
        h = C.c_void_p()
        if lib.orc_scene_create(C.byref(d), C.byref(h)) != 0:
            raise RuntimeError("orc_scene_create failed")
        self._h = h

    @classmethod
    def from_json(cls, path: str, texture_dir: str = "../../Textures") -> "OracleScene":
        return cls(*scene_io.load_scene(path, texture_dir))

    @classmethod
    def from_dict(cls, scene: dict, texture_dir: str = "../../Textures") -> "OracleScene":
        return cls(*scene_io.scene_arrays(scene, texture_dir))

    def __del__(self):
        try:
            if self._h:
                _load().orc_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def shape_order(self) -> np.ndarray:
        out = np.zeros(self.n_shapes, dtype=np.int32)
        assert _load().orc_scene_shape_order(self._h, out.ctypes.data, self.n_shapes) == 0
        return out

    def dump_bvh(self) -> list:
        buf = (_NodeDump * max(1, 2 * self.n_shapes))()
        n = _load().orc_scene_dump_bvh(self._h, buf, 2 * self.n_shapes)
        return [(b.is_leaf, tuple(b.lo), tuple(b.hi), [b.prims[k] for k in range(b.count)]) for b in buf[:n]]

    def render(self, use_bvh=False, n_samples_sqrt=4, light_samples=1, max_depth=10, seed=1, fixed_time=-1.0,
               rows=None, threads=None, cols=None):
        """Returns dict(rgb, ids, t, linear, rays=(primary, shadow, secondary)). rows / cols = (first, one past
        the last): only that window of the frame is rendered (the other pixels of the full-size outputs stay 0 / -1)."""
        p = _Params(int(bool(use_bvh)), int(n_samples_sqrt), int(light_samples), int(max_depth), int(seed), float(fixed_time),
                    rows[0] if rows else 0, rows[1] if rows else 0, threads or (os.cpu_count() or 1),
                    cols[0] if cols else 0, cols[1] if cols else 0)
        h, w = self.height, self.width
        rgb = np.zeros((h, w, 3), dtype=np.uint8)
        ids = np.full((h, w), -1, dtype=np.int32)
        t = np.zeros((h, w), dtype=np.float32)
        lin = np.zeros((h, w, 3), dtype=np.float32)
        rays = np.zeros(3, dtype=np.uint64)
        rc = _load().orc_render(self._h, C.byref(p), rgb.ctypes.data, ids.ctypes.data, t.ctypes.data, lin.ctypes.data, rays.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"orc_render failed: {rc}")
        return {"rgb": rgb, "ids": ids, "t": t, "linear": lin, "rays": tuple(int(x) for x in rays)}


class RefDriver:
    """oracle/_ref/ref_driver: the unmodified reference behind a small driver (ref_driver.cpp)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)

    @staticmethod
    def _run(args, cwd=None):
        r = subprocess.run([REF_DRIVER] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, cwd=cwd, check=True)
        last = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
        return json.loads(last[-1]) if last else {}

    @classmethod
    def ids(cls, scene_path, use_bvh=True, time=0.0, tmp="/tmp", cwd=None, rows=None, cols=None):
        """(ids, t, info) of the rows [rows[0], rows[1]) (default: all): arrays of shape (rows, width); with
        cols only those columns are computed (the others stay 0)."""
        out = os.path.join(tmp, f"ref_ids_{os.getpid()}.bin")
        args = ["--scene", scene_path, "--mode", "ids", "--bvh", str(int(use_bvh)), "--time", str(time), "--out-ids", out]
        if rows:
            args += ["--rows", str(rows[0]), str(rows[1])]
        if cols:
            args += ["--cols", str(cols[0]), str(cols[1])]
        info = cls._run(args, cwd)
        raw = np.fromfile(out, dtype=np.int32)
        os.remove(out)
        w, h = int(raw[0]), int(raw[1])
        ids = raw[4:4 + w * h].reshape(h, w).copy()
        t = raw[4 + w * h:4 + 2 * w * h].view(np.float32).reshape(h, w).copy()
        return ids, t, info

    @classmethod
    def render(cls, scene_path, use_bvh=True, n_samples_sqrt=1, light_samples=1, max_depth=10, seed=1, rows=None,
               tmp="/tmp", cwd=None, cols=None):
        out = os.path.join(tmp, f"ref_raw_{os.getpid()}.bin")
        args = ["--scene", scene_path, "--mode", "render", "--bvh", str(int(use_bvh)), "--s", str(n_samples_sqrt),
                "--light-samples", str(light_samples), "--depth", str(max_depth), "--seed", str(seed), "--out-raw", out]
        if rows:
            args += ["--rows", str(rows[0]), str(rows[1])]
        if cols:
            args += ["--cols", str(cols[0]), str(cols[1])]
        info = cls._run(args, cwd)
        raw = np.fromfile(out, dtype=np.uint8)
        os.remove(out)
        hdr = raw[:16].view(np.int32)
        w, h = int(hdr[0]), int(hdr[1])
        rgb = raw[16:16 + w * h * 3].reshape(h, w, 3).copy()
        lin = raw[16 + w * h * 3:16 + w * h * 3 + w * h * 12].view(np.float32).reshape(h, w, 3).copy()
        return rgb, lin, info

    @classmethod
    def bvh(cls, scene_path, tmp="/tmp"):
        out = os.path.join(tmp, f"ref_bvh_{os.getpid()}.txt")
        cls._run(["--scene", scene_path, "--mode", "bvh", "--out-bvh", out])
        nodes = []
        with open(out) as f:
            lines = f.read().splitlines()[1:]
        os.remove(out)
        for line in lines:
            tok = line.split()
            box = [np.float32(float.fromhex(x)) for x in tok[1:7]]
            prims = [int(x) for x in tok[8:]] if tok[0] == "L" else []
            nodes.append((int(tok[0] == "L"), tuple(float(b) for b in box[:3]), tuple(float(b) for b in box[3:]), prims))
        return nodes
