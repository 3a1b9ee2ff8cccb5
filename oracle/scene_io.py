"""oracle/scene_io.py -- TEST INFRASTRUCTURE: scene.json -> constructor-level arrays, in Python.

An independent restatement of the reference loader (Code/json_loader.cpp:30-338, camera.cpp:14-58)
on top of Python's own ``json`` module, so that the oracle does not share the product's C++ JSON
reader. Numbers become IEEE doubles in Python and are narrowed with ``np.float32`` exactly like
nlohmann's ``get<float>()`` (static_cast<float>).
"""
from __future__ import annotations

import json
import os

import numpy as np

SHAPE_DTYPE = np.dtype([
    ("type", np.int32), ("material", np.int32),
    ("translation", np.float32, 3), ("rotation", np.float32, 3), ("scale", np.float32, 3),
    ("velocity", np.float32, 3), ("corners", np.float32, 12),
])
MATERIAL_DTYPE = np.dtype([("f", np.float32, 14), ("texture", np.int32)])
LIGHT_DTYPE = np.dtype([("location", np.float32, 3), ("color", np.float32, 3), ("intensity", np.float32), ("radius", np.float32)])

F = np.float32


def _f3(v):
    if not isinstance(v, list) or len(v) != 3:
        raise ValueError("expected 3 numbers")
    return [F(x) for x in v]


def read_ppm_p3(path):
    """Image::read (image.cpp:86-133). Returns (H, W, 3) uint8 or None."""
    if not os.path.exists(path):
        return None
    with open(path) as f:
        tokens = []
        for line in f:
            if line.lstrip().startswith("#"):
                continue
            tokens.extend(line.split())
    if not tokens or tokens[0] != "P3":
        return None
    w, h = int(tokens[1]), int(tokens[2])
    vals = np.array([int(t) for t in tokens[4:4 + w * h * 3]], dtype=np.int64)
    out = np.zeros(w * h * 3, dtype=np.uint8)
    out[: len(vals)] = np.clip(vals, 0, 255)
    return out.reshape(h, w, 3)


class _Tables:
    def __init__(self, texture_dir):
        self.texture_dir = texture_dir
        self.materials = []
        self.mat_index = {}
        self.textures = []
        self.tex_index = {}
        self.by_identity = {}  # id(material dict) -> index: generated scenes share one dict per palette entry

    def texture(self, name):
        if len(name) < 3:
            return -1
        path = os.path.join(self.texture_dir, name[:-3] + "ppm")  # json_loader.cpp:78-80
        if path not in self.tex_index:
            img = read_ppm_p3(path)
            if img is None or img.shape[1] == 0:
                self.tex_index[path] = -1
            else:
                self.tex_index[path] = len(self.textures)
                self.textures.append(img)
        return self.tex_index[path]

    def material(self, mj):
        if mj is not None and id(mj) in self.by_identity:
            return self.by_identity[id(mj)][1]
        try:
            idx = self._material(mj)
        except (KeyError, ValueError, TypeError):  # json_loader.cpp:91-95: a material that fails to parse is the default one
            idx = self._material(None)
        if mj is not None:
            self.by_identity[id(mj)] = (mj, idx)  # keeps mj alive, so the id stays unique
        return idx

    def _material(self, mj):
        # defaults of Material (material.hpp:52-70) when the block is absent
        if mj is None:
            vals = [F(0.8)] * 3 + [F(1.0)] * 3 + [F(0.1), F(0.9), F(0.3), F(20.0), F(0.0), F(0.0), F(0.0), F(1.0)]
            tex = -1
        else:
            d = _f3(mj["diffuse_color"]) if "diffuse_color" in mj else [F(0.8)] * 3
            s = _f3(mj["specular_color"]) if "specular_color" in mj else [F(1.0)] * 3
            ka = F(mj.get("k_ambient", 0.1))
            kd = F(mj.get("k_diffuse", 0.6))
            ks = F(mj.get("k_specular", 0.6))
            rough = max(F(0.001), F(mj.get("roughness", 0.001)))
            r = max(F(0.001), min(F(1.0), rough))
            shininess = F(5.0) / (r * r)  # json_loader.cpp:56-61
            roughness = F(mj.get("roughness", 0.0))
            vals = d + s + [ka, kd, ks, F(shininess), roughness, F(mj.get("reflectivity", 0.0)),
                            F(mj.get("transparency", 0.0)), F(mj.get("refractive_index", 1.0))]
            tex = -1
            tf = mj.get("texture_file", "")
            if isinstance(tf, str) and tf:
                tex = self.texture(tf)
        key = (tuple(float(v) for v in vals), tex)
        if key not in self.mat_index:
            self.mat_index[key] = len(self.materials)
            self.materials.append((vals, tex))
        return self.mat_index[key]


def scene_arrays(scene: dict, texture_dir: str = "../../Textures"):
    """Returns (camera dict, lights, materials, shapes, textures) at constructor level."""
    cj = scene["cameras"][0]
    camera = {
        "location": _f3(cj["location"]), "gaze": _f3(cj["gaze_vector"]), "up": _f3(cj["up_vector"]),
        "focal_length": F(cj["focal_length"]),
        "sensor_width": int(cj["sensor_width"]), "sensor_height": int(cj["sensor_height"]),  # get<int>(): truncation
        "aperture": F(cj.get("aperture", 0.0)), "focus_dist": F(cj.get("focus_dist", 10.0)),
        "res_x": int(scene["render"]["resolution_x"]), "res_y": int(scene["render"]["resolution_y"]),
    }
    lights = []
    for lj in scene.get("lights", []) if isinstance(scene.get("lights", []), list) else []:
        if not isinstance(lj, dict) or not all(k in lj for k in ("location", "color", "intensity")):
            continue
        if F(lj["intensity"]) <= 0:
            continue
        lights.append((_f3(lj["location"]), _f3(lj["color"]), F(lj["intensity"]), F(lj.get("radius", 0.0))))
    lights_a = np.zeros(len(lights), dtype=LIGHT_DTYPE)
    for i, (loc, col, inten, rad) in enumerate(lights):
        lights_a[i] = (loc, col, inten, rad)

    tabs = _Tables(texture_dir)
    shapes = []
    zero = [F(0)] * 3
    for sj in scene.get("spheres", []):
        if not isinstance(sj, dict):
            continue
        try:
            t = _f3(sj["location"])
            r = _f3(sj["rotation"]) if "rotation" in sj else list(zero)
            if isinstance(sj.get("scale"), list):
                sc = _f3(sj["scale"])
            elif "radius" in sj:
                sc = [F(sj["radius"])] * 3
            else:
                sc = [F(1)] * 3
            mat = tabs.material(sj.get("material"))
            vel = _f3(sj["velocity"]) if "velocity" in sj else list(zero)
            vel = [v / F(5) for v in vel]  # json_loader.cpp:221-223
            shapes.append((0, mat, t, r, sc, vel, [F(0)] * 12))
        except (KeyError, ValueError, TypeError):
            continue
    for cj2 in scene.get("cubes", []):
        if not isinstance(cj2, dict) or "translation" not in cj2 or "rotation" not in cj2:
            continue
        try:
            sc = [F(1)] * 3
            if "scale" in cj2:
                sc = _f3(cj2["scale"]) if isinstance(cj2["scale"], list) else [F(cj2["scale"])] * 3
            mat = tabs.material(cj2.get("material"))
            shapes.append((1, mat, _f3(cj2["translation"]), _f3(cj2["rotation"]), sc, list(zero), [F(0)] * 12))
        except (KeyError, ValueError, TypeError):
            continue
    for rj in scene.get("rectangles", []):
        if not isinstance(rj, dict):
            continue
        try:
            t, r, sc = _f3(rj["translation"]), _f3(rj["rotation"]), _f3(rj["scale"])
            mat = tabs.material(rj.get("material"))
            shapes.append((2, mat, t, r, sc, list(zero), [F(0)] * 12))
        except (KeyError, ValueError, TypeError):
            continue
    # Bulk path for big generated scenes (millions of well-formed quads): corners converted by numpy in one
    # go, same narrowing double -> float32. Anything irregular falls through to the per-entry loop below.
    planes = scene.get("planes", [])
    plane_block = None
    if len(planes) > 10000:
        try:
            corners = np.asarray([pj["corners"] for pj in planes], dtype=np.float64)
            if corners.shape == (len(planes), 4, 3) and all(type(c[0][0]) in (float, int) for c in (planes[0]["corners"], planes[-1]["corners"])):
                plane_block = np.zeros(len(planes), dtype=SHAPE_DTYPE)
                plane_block["type"] = 3
                plane_block["material"] = [tabs.material(pj.get("material")) for pj in planes]
                plane_block["corners"] = corners.reshape(len(planes), 12).astype(np.float32)
                planes = []
        except (KeyError, ValueError, TypeError):
            plane_block = None
            planes = scene.get("planes", [])
    for pj in planes:
        if not isinstance(pj, dict) or not isinstance(pj.get("corners"), list) or len(pj["corners"]) != 4:
            continue
        try:
            corners = [c for k in range(4) for c in _f3(pj["corners"][k])]
            mat = tabs.material(pj.get("material"))
            shapes.append((3, mat, list(zero), list(zero), list(zero), list(zero), corners))
        except (KeyError, ValueError, TypeError):
            continue
    shapes_a = np.zeros(len(shapes), dtype=SHAPE_DTYPE)
    for i, s in enumerate(shapes):
        shapes_a[i] = s
    if plane_block is not None:
        shapes_a = np.concatenate([shapes_a, plane_block])
    mats_a = np.zeros(max(1, len(tabs.materials)), dtype=MATERIAL_DTYPE)
    if not tabs.materials:
        tabs.material(None)
    for i, (vals, tex) in enumerate(tabs.materials):
        mats_a[i] = (vals, tex)
    return camera, lights_a, mats_a, shapes_a, tabs.textures


def load_scene(path: str, texture_dir: str = "../../Textures"):
    with open(path) as f:
        return scene_arrays(json.load(f), texture_dir)
