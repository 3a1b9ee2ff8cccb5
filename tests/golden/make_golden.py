"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/ref_driver):
    python tests/golden/make_golden.py
Every fixture = a scene in the reference's JSON schema (*.json, minified) + the reference's own
outputs for it (*.npz): primary-ray hit IDs and distances, 8-bit image, linear float image, and
for stochastic scenes a 4096-spp reference render. The GPU box has no /root/reference; tests
there use these files.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import RefDriver  # noqa: E402
from ray_tracying_b200 import scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save_scene(name, scene):
    path = os.path.join(OUT, name + ".json")
    scenes.write_scene(scene, path)
    return path


def deterministic(name, scene, depth=10, cwd=None):
    path = save_scene(name, scene)
    out = {}
    for bvh in (1, 0):
        ids, t, _ = RefDriver.ids(path, use_bvh=bool(bvh), cwd=cwd)
        rgb, lin, _ = RefDriver.render(path, use_bvh=bool(bvh), n_samples_sqrt=1, max_depth=depth, cwd=cwd)
        out[f"ids_bvh{bvh}"] = ids
        out[f"t_bvh{bvh}"] = t
        out[f"rgb_bvh{bvh}"] = rgb
        out[f"linear_bvh{bvh}"] = lin
    out["depth"] = np.int32(depth)
    out["bvh_dump"] = np.array(json.dumps(RefDriver.bvh(path)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def stochastic(name, scene, s=64, light_samples=1, depth=10, seed=7):
    path = save_scene(name, scene)
    ids, t, _ = RefDriver.ids(path, use_bvh=True)
    rgb, lin, info = RefDriver.render(path, use_bvh=True, n_samples_sqrt=s, light_samples=light_samples, max_depth=depth, seed=seed)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ids_bvh1=ids, t_bvh1=t, ref_rgb=rgb, ref_linear=lin,
                        ref_spp=np.int32(s * s), light_samples=np.int32(light_samples), depth=np.int32(depth))
    print(name, "spp", s * s, "seconds", info.get("seconds"))


def numerics_edge_scene():
    """Cases that stress the fast traversal's exactness machinery: an axis-aligned camera with an ODD
    width (the centre column has d.x == 0 exactly -> the reference's 'parallel' rule, and its
    neighbours have |d.x| just above 1e-6), a light exactly above the origin (vertical shadow rays),
    degenerate quads (repeated corner: the reference's inside test accepts a whole line), a non-planar
    quad, sub-pixel and strongly anisotropic spheres (discriminant cancellation), axis-aligned cubes
    (box == shape) and glass / mirror materials so that secondary rays start on all of them."""
    sc = scenes.mixed_scene(0, seed=2, resolution=(161, 91), extent=6.0, height=1.0, fractions=(1, 0, 0, 0))
    sc["cameras"] = [scenes.camera_block((0.0, -9.0, 1.5), (0.0, 0.0, 1.5), focal_length=30.0)]
    sc["lights"] = [{"location": [0.0, 0.0, 6.0], "intensity": 900.0, "color": [1.0, 1.0, 1.0], "radius": 0.0},
                    {"location": [4.0, -5.0, 3.0], "intensity": 700.0, "color": [1.0, 0.9, 0.8], "radius": 0.0}]
    m = scenes.material_block
    glass = m(diffuse=(0.95, 0.95, 1.0), roughness=0.0, reflectivity=0.1, transparency=0.8, refractive_index=1.5)
    mirror = m(diffuse=(0.9, 0.9, 0.9), roughness=0.0, reflectivity=0.6)
    red, green, blue = m(diffuse=(0.9, 0.2, 0.2)), m(diffuse=(0.2, 0.9, 0.2)), m(diffuse=(0.2, 0.3, 0.9))
    sc["rectangles"] = [
        {"translation": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0], "scale": [30.0, 30.0, 1.0], "material": mirror},
        {"translation": [0.0, 3.0, 1.5], "rotation": [1.5707963, 0.0, 0.0], "scale": [2.0, 1.0, 1.0], "material": blue},
    ]
    sc["spheres"] = [
        {"location": [0.0, 0.0, 1.5], "rotation": [0.0, 0.0, 0.0], "scale": [1.0, 1.0, 1.0], "material": glass},
        {"location": [0.3, 40.0, 3.0], "rotation": [0.0, 0.0, 0.0], "scale": [0.004, 0.004, 0.004], "material": red},
        {"location": [-0.8, 35.0, 2.0], "rotation": [0.3, 0.2, 0.1], "scale": [0.02, 0.02, 0.02], "material": red},
        {"location": [2.5, 1.0, 1.0], "rotation": [0.4, 0.9, 0.2], "scale": [1.2, 0.9, 0.01], "material": green},
        {"location": [-2.5, 0.5, 0.8], "rotation": [0.0, 0.0, 0.0], "scale": [0.03, 0.8, 0.8], "material": mirror},
        {"location": [-4.0, 2.0, 2.5], "rotation": [0.0, 0.0, 0.0], "scale": [0.5, 0.5, 0.5], "material": blue},
    ]
    sc["cubes"] = [
        {"translation": [0.0, 5.0, 1.0], "rotation": [0.0, 0.0, 0.0], "scale": [2.0, 2.0, 2.0], "material": mirror},
        {"translation": [3.0, -2.0, 0.5], "rotation": [0.0, 0.0, 0.0], "scale": [1.0, 1.0, 1.0], "material": green},
        {"translation": [-3.0, -3.0, 0.25], "rotation": [0.0, 0.0, 0.7853982], "scale": [0.5, 3.0, 0.5], "material": red},
        {"translation": [1.2, -4.0, 1.5], "rotation": [0.5, 0.3, 0.1], "scale": [0.02, 0.6, 0.6], "material": glass},
    ]
    sc["planes"] = [
        # proper quad, triangle with a repeated corner (c3 == c2), one with c1 == c0 (no valid normal)
        {"corners": [[-5, 6, 0.5], [-3, 6, 0.5], [-5, 6, 2.5], [-3, 6, 2.5]], "material": green},
        {"corners": [[3, 6, 0.5], [5, 6, 0.5], [4, 6, 2.5], [4, 6, 2.5]], "material": red},
        {"corners": [[1, 2, 3], [1, 2, 3], [2, 2, 3], [2, 3, 3]], "material": red},
        # non-planar quad (corner 3 lifted off the plane of corners 0..2), sliver triangle
        {"corners": [[-1.5, -2, 0.2], [-0.5, -2, 0.2], [-1.5, -1, 0.2], [-0.5, -1, 0.9]], "material": blue},
        {"corners": [[1.5, -3, 0.3], [3.5, -3, 0.3], [2.5, -2.999, 0.3], [2.5, -2.999, 0.3]], "material": blue},
        {"corners": [[-6, -1, 0.1], [-5.9999, -1, 0.1], [-6, 1, 2.0], [-5.9999, 1, 2.0]], "material": green},
    ]
    return sc


def main():
    if not RefDriver.available():
        raise SystemExit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "numerics_edge":
        deterministic("numerics_edge", numerics_edge_scene(), depth=4)
        return

    # (0) the reference's own scene (ASCII/scene.json), minified; glossy, so only IDs/t are deterministic
    with open("/root/reference/ASCII/scene.json") as f:
        ascii_scene = json.load(f)
    save_scene("ascii_scene", ascii_scene)
    small = scenes.set_resolution(ascii_scene, 480, 270)
    path = save_scene("ascii_scene_480", small)
    out = {}
    for bvh in (1, 0):
        ids, t, _ = RefDriver.ids(path, use_bvh=bool(bvh))
        out[f"ids_bvh{bvh}"], out[f"t_bvh{bvh}"] = ids, t
    out["bvh_dump"] = np.array(json.dumps(RefDriver.bvh(path)))
    np.savez_compressed(os.path.join(OUT, "ascii_scene_480.npz"), **out)
    os.remove(path)  # the 480x270 variant is derived in the tests by changing the resolution
    stochastic("ascii_scene_96", scenes.set_resolution(ascii_scene, 96, 54), s=64)
    os.remove(os.path.join(OUT, "ascii_scene_96.json"))

    # (1) deterministic mixed scene: all four primitives, mirrors, glass, two point lights
    deterministic("mixed_400", scenes.mixed_scene(400, seed=3, resolution=(320, 180)))
    deterministic("mixed_400_depth5", scenes.mixed_scene(400, seed=3, resolution=(160, 90)), depth=5)

    # (2) edge cases
    deterministic("empty", {**scenes.mixed_scene(0, seed=1, resolution=(32, 18), fractions=(1, 0, 0, 0)), "rectangles": []})
    deterministic("few_3", scenes.mixed_scene(3, seed=5, resolution=(96, 54), extent=2.0, height=1.0, fractions=(1, 1, 0, 1)))
    deterministic("few_5", scenes.mixed_scene(5, seed=6, resolution=(96, 54), extent=2.0, height=1.0, fractions=(1, 1, 0, 1)))
    # coincident shapes (ties in t -> first in leaf order wins), a degenerate plane, an axis-aligned view
    tie = scenes.mixed_scene(0, seed=2, resolution=(96, 54), extent=2.0, height=1.0, fractions=(1, 0, 0, 0))
    tie["cameras"] = [scenes.camera_block((0.0, -6.0, 1.0), (0.0, 0.0, 1.0))]
    red, green = scenes.material_block(diffuse=(0.9, 0.1, 0.1)), scenes.material_block(diffuse=(0.1, 0.9, 0.1))
    tie["cubes"] = [
        {"translation": [0.0, 0.0, 1.0], "rotation": [0.0, 0.0, 0.0], "scale": [1.5, 1.5, 1.5], "material": red},
        {"translation": [0.0, 0.0, 1.0], "rotation": [0.0, 0.0, 0.0], "scale": [1.5, 1.5, 1.5], "material": green},
        {"translation": [2.0, 0.0, 1.0], "rotation": [0.0, 0.0, 0.0], "scale": [1.0, 1.0, 1.0], "material": green},
    ]
    tie["spheres"] = [
        {"location": [-2.0, 0.0, 1.0], "scale": [0.7, 0.7, 0.7], "material": red},
        {"location": [-2.0, 0.0, 1.0], "radius": 0.7, "material": green},
    ]
    tie["planes"] = [
        {"corners": [[-1, 1, 0.2], [1, 1, 0.2], [-1, 1, 2.2], [1, 1, 2.2]], "material": green},
        {"corners": [[-1, 1, 0.2], [1, 1, 0.2], [-1, 1, 2.2], [1, 1, 2.2]], "material": red},
        {"corners": [[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]]},
        {"corners": [[-3, -1, 3], [-2, -1, 3], [-3, -1, 3], [-2, -1, 3]]},
    ]
    deterministic("ties_axis_aligned", tie)
    deterministic("numerics_edge", numerics_edge_scene(), depth=4)

    # (3) textures: the reference resolves "x.jpg" to ../../Textures/x.ppm relative to its cwd
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "Textures"))
    os.makedirs(os.path.join(tmp, "Code", "build"))
    scenes.checker_texture(os.path.join(tmp, "Textures", "checker.ppm"), n=32, tiles=4, seed=11)
    shutil.copy(os.path.join(tmp, "Textures", "checker.ppm"), os.path.join(OUT, "checker.ppm"))
    tex = scenes.mixed_scene(40, seed=9, resolution=(192, 108), extent=3.0, height=1.5, texture_file="checker.jpg")
    for key in ("spheres", "cubes", "rectangles", "planes"):
        for i, sh in enumerate(tex.get(key, [])):
            if i % 2 == 0:
                sh["material"] = dict(sh["material"], texture_file="checker.jpg")
    deterministic("textured_40", tex, cwd=os.path.join(tmp, "Code", "build"))
    shutil.rmtree(tmp)

    # (4) stochastic effects, 4096-spp reference renders at 64x36
    res = (64, 36)
    stochastic("soft_shadows", scenes.mixed_scene(60, seed=21, resolution=res, extent=3.0, height=1.5, light_radius=0.6,
                                                  glass=False, mirror=False), s=64, light_samples=4)
    stochastic("glossy", scenes.mixed_scene(60, seed=22, resolution=res, extent=3.0, height=1.5, glossy=True, mirror=False,
                                            glass=False), s=64)
    stochastic("dof", scenes.mixed_scene(60, seed=23, resolution=res, extent=3.0, height=1.5, aperture=0.35, glass=False,
                                         mirror=False), s=64)
    stochastic("motion_blur", scenes.mixed_scene(60, seed=24, resolution=res, extent=3.0, height=1.5, moving_fraction=0.6,
                                                 fractions=(0.7, 0.1, 0.1, 0.1), glass=False, mirror=False), s=64)
    stochastic("antialias", scenes.mixed_scene(60, seed=25, resolution=res, extent=3.0, height=1.5), s=64)


if __name__ == "__main__":
    main()
