"""GPU property tests (-m gpu) of the exactness machinery itself, below the level of whole frames:

* the conservative slab test (rt_device.cuh: wide_child_test) never rejects a box that the reference's
  exact AABB::intersect (shapes.cpp:55-72) accepts, and "surely passes" implies the exact test passes --
  on hundreds of millions of random (ray, box) pairs including near-axis-parallel rays, huge coordinates
  and rays grazing faces / edges / corners to a few ulps;
* a primitive's culling box (bvh.cpp: cull_pad) is never missed by a ray that the exact intersection
  routine reports as a hit -- for every primitive of several scenes, from near and far origins.
"""
import os

import pytest

from conftest import GOLDEN, scene_file

pytestmark = pytest.mark.gpu


def test_conservative_box_test_never_contradicts_the_exact_one(rt):
    total = {"tests": 0, "exact": 0, "conservative": 0, "surely": 0, "skipped": 0}
    for seed in (1, 2, 3, 4):
        r = rt.selftest_boxes(50_000_000, seed=seed)
        assert r["violations_exact_not_conservative"] == 0, r
        assert r["violations_surely_not_exact"] == 0, r
        for k in total:
            total[k] += r[k]
    # the test must not be vacuous: many boxes are hit, many missed; a third of the pairs graze the box to
    # a few ulps (those are the ones left to the exact test), the rest is decided by "surely"
    assert total["tests"] > 150_000_000
    assert 0.1 < total["exact"] / total["tests"] < 0.9
    assert total["conservative"] >= total["exact"] >= total["surely"] > 0.5 * total["exact"]
    assert total["conservative"] < 1.35 * total["exact"], "the conservative test should stay tight"


@pytest.mark.parametrize("name", ["mixed_400", "numerics_edge", "ties_axis_aligned", "motion_blur", "ascii_scene"])
def test_culling_boxes_contain_every_exact_hit_golden_scenes(rt, name):
    scene = rt.Scene.from_json(os.path.join(GOLDEN, name + ".json"), GOLDEN)
    r = scene.selftest_cull(rays_per_primitive=20000, seed=7)
    assert r["violations"] == 0, r
    assert r["tests"] > 0 and 0 < r["hits"] < r["tests"] and r["passes"] >= r["hits"]


def test_culling_boxes_contain_every_exact_hit_large_scene(rt, tmp_path):
    """Far origins relative to small primitives (the distance-squared rounding term of the sphere test)."""
    from ray_tracying_b200 import scenes
    sc = scenes.mixed_scene(n_shapes=20000, seed=5, resolution=(64, 36), extent=200.0, height=20.0, fill=0.002)
    scene = rt.Scene.from_json(scene_file(tmp_path, sc), GOLDEN)
    r = scene.selftest_cull(rays_per_primitive=2000, seed=3)
    assert r["violations"] == 0, r
    assert r["hits"] > 1_000_000


def test_traversal_ceiling_and_launch_accounting(rt):
    """rt_traversal_peak (the roofline denominator of bench.py) returns a plausible, repeatable rate for both query
    flavours, and rt_scene_launch_count adds exactly what each frame reports."""
    import os
    from conftest import GOLDEN
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "mixed_400.json"), GOLDEN)
    a, ms_a = scene.traversal_peak(any_hit=False, steps=2048, repeats=3)
    b, _ = scene.traversal_peak(any_hit=True, steps=2048, repeats=3)
    a2, _ = scene.traversal_peak(any_hit=False, steps=2048, repeats=3)
    assert 2e10 < a < 5e12 and 2e10 < b < 5e12 and ms_a > 0
    assert abs(a - a2) / a < 0.1
    n0 = scene.launch_count()
    st = scene.render(use_bvh=True, n_samples_sqrt=1, fixed_time=0.0, max_depth=3)[3]
    assert scene.launch_count() - n0 == st.launches == 1 + 1 + 4 * 4 + 1 + 1
