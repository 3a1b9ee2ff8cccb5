"""The benchmark workloads = BASELINE.json's configs, shared by bench.py and the parity tests so that what is
measured is exactly what is checked against the reference. Pure Python + numpy: importing this module (or
``ray_tracying_b200.scenes``) does not load librt_b200.so.

    mixed100k : 100k-shape mixed scene (spheres/ellipsoids, cubes incl. rod-like ones, rectangles,
                plane quads = 2 triangles each), 1920x1080, 1 spp, Whitted depth 5       [configs[1]]
    soup1m    : 1M-triangle soup (500k plane quads), 1080p, 64 spp, 16-sample area light  [configs[2]]
    glossy250k: 250k-triangle glossy scene, 3840x2160, 100 spp, depth 8                   [configs[3]]
    dof4m     : 4M-triangle scene, thin lens + motion blur, 3840x2160, 256 spp            [configs[4]]
    ascii     : the reference's own ASCII/scene.json, 1 spp                               [configs[0]]

The reference has no triangle primitive: a "triangle" scene is made of plane quads (shapes.cpp:485-494
tests a quad as two triangles), see scenes.py.
"""
from __future__ import annotations

import os
import sys
import time

from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.environ.get("RT_BENCH_CACHE", "/tmp/rt_b200_bench")

WORKLOADS = {
    # name: generator (function of scenes.py, kwargs), render switches (the reference CLI's), description
    "mixed100k": dict(gen=("mixed_scene", dict(n_shapes=100000, seed=1, resolution=(1920, 1080), extent=30.0, height=6.0)),
                      render=dict(use_bvh=True, n_samples_sqrt=1, light_samples=1, max_depth=5),
                      desc="configs[1]: 100k-shape mixed scene, 1920x1080, 1 spp, Whitted depth 5"),
    "soup1m": dict(gen=("quad_soup", dict(n_triangles=1000000, seed=2, resolution=(1920, 1080), extent=40.0, height=8.0,
                                          light_radius=2.0, n_lights=1)),
                   render=dict(use_bvh=True, n_samples_sqrt=8, light_samples=16, max_depth=10),
                   desc="configs[2]: 1M-triangle soup (500k quads), 1920x1080, 64 spp, 16-sample area light"),
    "glossy250k": dict(gen=("quad_soup", dict(n_triangles=250000, seed=3, resolution=(3840, 2160), extent=25.0, height=6.0,
                                              glossy=True, n_lights=2)),
                       render=dict(use_bvh=True, n_samples_sqrt=10, light_samples=1, max_depth=8),
                       desc="configs[3]: 250k-triangle glossy scene, 3840x2160, 100 spp, depth 8"),
    "dof4m": dict(gen=("quad_soup", dict(n_triangles=4000000, seed=4, resolution=(3840, 2160), extent=60.0, height=10.0,
                                         aperture=0.8, n_moving_spheres=64, n_lights=2)),
                  render=dict(use_bvh=True, n_samples_sqrt=16, light_samples=1, max_depth=10),
                  desc="configs[4]: 4M-triangle scene, thin-lens DOF + motion blur, 3840x2160, 256 spp"),
    "ascii": dict(gen=("ascii", {}), render=dict(use_bvh=True, n_samples_sqrt=1, light_samples=1, max_depth=10),
                  desc="configs[0]: the reference's ASCII/scene.json, 1920x1080, 1 spp"),
}


def scene_dict(name: str) -> dict:
    """The workload's scene in the reference's scene.json schema (a dict)."""
    if name == "ascii":
        import json
        with open(os.path.join(ROOT, "tests", "golden", "ascii_scene.json")) as f:
            return json.load(f)
    fn, kw = WORKLOADS[name]["gen"]
    return getattr(scenes, fn)(**kw)


def scene_path_for(name: str, scene: dict | None = None) -> str:
    """Path of the workload's scene.json, generated into the cache directory on first use."""
    os.makedirs(CACHE, exist_ok=True)
    if name == "ascii":
        return os.path.join(ROOT, "tests", "golden", "ascii_scene.json")
    path = os.path.join(CACHE, name + ".json")
    if not os.path.exists(path):
        t0 = time.time()
        sc = scene if scene is not None else scene_dict(name)
        tmp = f"{path}.{os.getpid()}.tmp"
        scenes.write_scene(sc, tmp)
        os.replace(tmp, path)
        print(f"[workloads] generated {name}: {scenes.shape_count(sc)} shapes, {os.path.getsize(path) / 1e6:.1f} MB in {time.time() - t0:.1f}s",
              file=sys.stderr)
    return path
