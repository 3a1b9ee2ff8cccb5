"""Synthetic scenes in the reference's scene.json schema (reference Blend/exporter.py:182-270,
Code/json_loader.cpp:164-338). A generated dict can be dumped to a file and read by this
package (Scene.from_json), by the oracle and by the unmodified reference binary alike.

The reference has four primitives: spheres (ellipsoids, optionally moving), cubes, rectangles
and planes (a quad given by 4 corners, tested as two triangles, shapes.cpp:485-494). It has no
triangle or cylinder primitive, so "triangle" workloads are expressed as plane quads -- one quad
is two triangles -- and "cylinders" as elongated cubes / ellipsoids.
"""
from __future__ import annotations

import json
from typing import Optional

import numpy as np


def _r(a, nd=4):
    return np.round(np.asarray(a, dtype=np.float64), nd).tolist()


def camera_block(location, look_at, up=(0.0, 0.0, 1.0), focal_length=35.0, sensor=(36, 24), aperture=0.0,
                 focus_dist=None):
    loc = np.asarray(location, dtype=np.float64)
    tgt = np.asarray(look_at, dtype=np.float64)
    gaze = tgt - loc
    dist = float(np.linalg.norm(gaze))
    gaze = gaze / dist
    upv = np.asarray(up, dtype=np.float64)
    upv = upv - gaze * float(upv @ gaze)
    upv = upv / np.linalg.norm(upv)
    return {
        "location": _r(loc, 5), "gaze_vector": _r(gaze, 6), "up_vector": _r(upv, 6),
        "focal_length": float(focal_length), "sensor_width": float(sensor[0]), "sensor_height": float(sensor[1]),
        "aperture": float(aperture), "focus_dist": float(round(dist if focus_dist is None else focus_dist, 4)),
    }


def material_block(diffuse=(0.8, 0.8, 0.8), specular=(1.0, 1.0, 1.0), roughness=0.3, k_ambient=0.1, k_diffuse=0.7,
                   k_specular=0.4, reflectivity=0.0, transparency=0.0, refractive_index=1.0, texture_file=""):
    return {
        "diffuse_color": _r(diffuse, 3), "specular_color": _r(specular, 3), "roughness": round(float(roughness), 4),
        "k_ambient": k_ambient, "k_diffuse": k_diffuse, "k_specular": k_specular,
        "reflectivity": round(float(reflectivity), 3), "transparency": round(float(transparency), 3),
        "refractive_index": round(float(refractive_index), 3), "texture_file": texture_file,
    }


def _palette(rng, glossy: bool, glass: bool, mirror: bool):
    """A small material palette: matte colours, mirrors (roughness 0 unless glossy), glass."""
    mats = []
    for _ in range(6):
        mats.append(material_block(diffuse=rng.uniform(0.2, 0.95, 3), roughness=float(rng.uniform(0.15, 0.6))))
    if mirror:
        mats.append(material_block(diffuse=(0.9, 0.9, 0.9), roughness=0.0, reflectivity=0.6, k_specular=0.5))
        mats.append(material_block(diffuse=(0.8, 0.6, 0.3), roughness=0.0, reflectivity=0.3, k_specular=0.5))
    if glossy:
        mats.append(material_block(diffuse=(0.7, 0.8, 0.9), roughness=0.08, reflectivity=0.5))
        mats.append(material_block(diffuse=(0.9, 0.7, 0.6), roughness=0.2, reflectivity=0.4))
    if glass:
        mats.append(material_block(diffuse=(0.95, 0.95, 1.0), roughness=0.0, reflectivity=0.1, transparency=0.8,
                                   refractive_index=1.5))
    return mats


def mixed_scene(n_shapes: int, seed: int = 0, resolution=(1920, 1080), extent: float = 10.0, height: float = 4.0,
                fractions=(0.35, 0.25, 0.1, 0.3), glossy: bool = False, glass: bool = True, mirror: bool = True,
                light_radius: float = 0.0, n_lights: int = 2, aperture: float = 0.0, moving_fraction: float = 0.0,
                fill: float = 0.35, texture_file: str = "") -> dict:
    """Spheres/ellipsoids, cubes (some elongated), rectangles and plane quads scattered in a slab
    [-extent,extent]^2 x [0,height] over a floor rectangle. `fractions` = (spheres, cubes,
    rectangles, planes). Deterministic (no RNG in the renderer) unless glossy, light_radius,
    aperture or moving_fraction is set."""
    rng = np.random.default_rng(seed)
    counts = (np.asarray(fractions, dtype=np.float64) / np.sum(fractions) * n_shapes).astype(int)
    counts[0] += n_shapes - counts.sum()
    n_sph, n_cube, n_rect, n_plane = (int(c) for c in counts)
    volume = (2 * extent) ** 2 * height
    size = float((fill * volume / max(n_shapes, 1)) ** (1.0 / 3.0))  # typical object size
    mats = _palette(rng, glossy, glass, mirror)

    def positions(n):
        p = rng.uniform(-1.0, 1.0, (n, 3))
        p[:, 0] *= extent
        p[:, 1] *= extent
        p[:, 2] = (p[:, 2] * 0.5 + 0.5) * height + 0.6 * size
        return p

    def pick(n):
        return rng.integers(0, len(mats), n)

    scene = {"cameras": [camera_block((-1.6 * extent, -1.9 * extent, 1.4 * height + 0.35 * extent), (0.0, 0.0, 0.3 * height),
                                      focal_length=40.0, aperture=aperture)]}
    lights = []
    for k in range(n_lights):
        ang = 2.0 * np.pi * (k + 0.25) / max(n_lights, 1)
        loc = (0.7 * extent * np.cos(ang), 0.7 * extent * np.sin(ang), height + 0.6 * extent)
        d2 = float(np.dot(loc, loc))
        lights.append({"location": _r(loc, 3), "intensity": round(0.9 * (25.0 + 150.0 * d2) / 10.0, 1),
                       "color": [1.0, 1.0 - 0.1 * (k % 2), 1.0 - 0.05 * k], "radius": float(light_radius)})
    scene["lights"] = lights

    if n_sph:
        pos, rot = positions(n_sph), rng.uniform(0, 2 * np.pi, (n_sph, 3))
        scl = size * 0.5 * rng.uniform(0.5, 1.2, (n_sph, 1)) * rng.uniform(0.6, 1.4, (n_sph, 3))
        mi = pick(n_sph)
        moving = rng.uniform(0, 1, n_sph) < moving_fraction
        vel = rng.uniform(-1, 1, (n_sph, 3)) * size * 4.0
        pos_l, rot_l, scl_l, vel_l = _r(pos), _r(rot), _r(scl), _r(vel, 3)
        scene["spheres"] = [
            {"location": pos_l[i], "rotation": rot_l[i], "scale": scl_l[i],
             "velocity": vel_l[i] if moving[i] else [0.0, 0.0, 0.0], "material": mats[mi[i]]} for i in range(n_sph)]
    if n_cube:
        pos, rot = positions(n_cube), rng.uniform(0, 2 * np.pi, (n_cube, 3))
        scl = size * rng.uniform(0.4, 1.1, (n_cube, 3))
        rods = rng.uniform(0, 1, n_cube) < 0.3  # "cylinder-like" elongated boxes
        scl[rods, 2] *= 2.5
        scl[rods, :2] *= 0.5
        mi = pick(n_cube)
        pos_l, rot_l, scl_l = _r(pos), _r(rot), _r(scl)
        scene["cubes"] = [{"translation": pos_l[i], "rotation": rot_l[i], "scale": scl_l[i], "material": mats[mi[i]]}
                          for i in range(n_cube)]
    rects = [{"translation": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0], "scale": [2.6 * extent, 2.6 * extent, 1.0],
              # reflective + roughness > 0 means glossy (stochastic) in the reference, so the mirror floor is sharp
              "material": material_block(diffuse=(0.75, 0.75, 0.7), roughness=(0.25 if glossy else 0.0) if mirror else 0.5,
                                         reflectivity=0.15 if mirror else 0.0, texture_file=texture_file)}]
    if n_rect > 1:
        n = n_rect - 1
        pos, rot = positions(n), rng.uniform(0, 2 * np.pi, (n, 3))
        scl = np.concatenate([size * rng.uniform(0.6, 1.6, (n, 2)), np.ones((n, 1))], axis=1)
        mi = pick(n)
        pos_l, rot_l, scl_l = _r(pos), _r(rot), _r(scl)
        rects += [{"translation": pos_l[i], "rotation": rot_l[i], "scale": scl_l[i], "material": mats[mi[i]]} for i in range(n)]
    scene["rectangles"] = rects
    if n_plane:
        scene["planes"] = _quads(rng, positions(n_plane), size, mats, pick(n_plane))
    scene["render"] = {"resolution_x": int(resolution[0]), "resolution_y": int(resolution[1])}
    return scene


def _quads(rng, centres, size, mats, mat_index):
    """Planar convex quads (two-triangle strips): a random orthonormal frame (a, b) per quad."""
    n = len(centres)
    a = rng.normal(size=(n, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.normal(size=(n, 3))
    b -= a * np.sum(a * b, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    hw = size * rng.uniform(0.4, 1.0, (n, 1))
    hh = size * rng.uniform(0.4, 1.0, (n, 1))
    skew = rng.uniform(-0.3, 0.3, (n, 1)) * hw
    # Triangle-strip order: the reference tests triangles (c0,c1,c2) and (c1,c3,c2) against the
    # normal of (c0,c1,c2) (shapes.cpp:485-494), so c3 must be the corner diagonal to c0.
    c0 = centres - a * hw - b * hh
    c1 = centres + a * hw - b * hh
    c2 = centres - a * (hw - skew) + b * hh
    c3 = centres + a * (hw + skew) + b * hh
    corners = np.round(np.stack([c0, c1, c2, c3], axis=1), 4).tolist()
    if mats is None:
        return [{"corners": corners[i]} for i in range(n)]
    return [{"corners": corners[i], "material": mats[mat_index[i]]} for i in range(n)]


def quad_soup(n_triangles: int, seed: int = 0, resolution=(1920, 1080), extent: float = 10.0, height: float = 5.0,
              light_radius: float = 0.0, n_lights: int = 1, aperture: float = 0.0, glossy: bool = False,
              n_moving_spheres: int = 0, fill: float = 0.25, with_materials: bool = True) -> dict:
    """A "triangle soup" in the reference's vocabulary: n_triangles / 2 plane quads (each quad is
    tested as two triangles, shapes.cpp:485-494) over a floor, lit by point or area lights."""
    rng = np.random.default_rng(seed)
    n_quads = max(1, n_triangles // 2)
    volume = (2 * extent) ** 2 * height
    size = float((fill * volume / n_quads) ** (1.0 / 3.0))
    p = rng.uniform(-1.0, 1.0, (n_quads, 3))
    p[:, 0] *= extent
    p[:, 1] *= extent
    p[:, 2] = (p[:, 2] * 0.5 + 0.5) * height + size
    mats = _palette(rng, glossy, glass=False, mirror=not glossy) if with_materials else None
    mi = rng.integers(0, len(mats), n_quads) if mats else None
    scene = {"cameras": [camera_block((-1.5 * extent, -1.8 * extent, 1.5 * height + 0.3 * extent), (0.0, 0.0, 0.35 * height),
                                      focal_length=40.0, aperture=aperture)]}
    lights = []
    for k in range(n_lights):
        ang = 2.0 * np.pi * (k + 0.4) / max(n_lights, 1)
        loc = (0.6 * extent * np.cos(ang), 0.6 * extent * np.sin(ang), height + 0.7 * extent)
        d2 = float(np.dot(loc, loc))
        lights.append({"location": _r(loc, 3), "intensity": round(1.1 * (25.0 + 150.0 * d2) / 10.0, 1),
                       "color": [1.0, 1.0, 1.0], "radius": float(light_radius)})
    scene["lights"] = lights
    if n_moving_spheres:
        pos = rng.uniform(-0.6, 0.6, (n_moving_spheres, 3)) * np.array([extent, extent, 0.0]) + np.array([0, 0, height + 1.0])
        vel = rng.uniform(-1, 1, (n_moving_spheres, 3)) * np.array([6.0, 6.0, 1.0])
        scene["spheres"] = [{"location": _r(pos[i]), "rotation": [0.0, 0.0, 0.0], "scale": [0.5, 0.5, 0.5],
                             "velocity": _r(vel[i], 3), "material": material_block(diffuse=(0.9, 0.3, 0.2))}
                            for i in range(n_moving_spheres)]
    scene["rectangles"] = [{"translation": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0],
                            "scale": [2.6 * extent, 2.6 * extent, 1.0],
                            "material": material_block(diffuse=(0.7, 0.7, 0.7), roughness=0.5)}]
    scene["planes"] = _quads(rng, p, size, mats, mi)
    scene["render"] = {"resolution_x": int(resolution[0]), "resolution_y": int(resolution[1])}
    return scene


def write_scene(scene: dict, path: str) -> str:
    # json.dumps runs the C encoder in one shot (json.dump iterates in Python: 6x slower on a 170 MB scene)
    with open(path, "w") as f:
        f.write(json.dumps(scene, separators=(",", ":")))
    return path


def shape_count(scene: dict) -> int:
    return sum(len(scene.get(k, [])) for k in ("spheres", "cubes", "rectangles", "planes"))


def set_resolution(scene: dict, width: int, height: int) -> dict:
    out = dict(scene)
    out["render"] = {"resolution_x": int(width), "resolution_y": int(height)}
    return out


def checker_texture(path: str, n: int = 64, tiles: int = 8, seed: Optional[int] = None) -> str:
    """Writes a small P3 PPM checker texture (the reference reads textures as P3, image.cpp:86-133)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    mask = ((xx * tiles // n) + (yy * tiles // n)) % 2
    img = np.where(mask[..., None] == 1, np.array([230, 230, 210]), np.array([60, 80, 140])).astype(np.int64)
    if seed is not None:
        img = np.clip(img + rng.integers(-20, 20, img.shape), 0, 255)
    with open(path, "w") as f:
        f.write(f"P3\n{n} {n}\n255\n")
        for row in img:
            f.write("  ".join(" ".join(str(int(v)) for v in px) for px in row) + "\n")
    return path
