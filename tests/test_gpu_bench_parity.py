"""GPU parity at the BENCHMARK shapes (-m gpu): every workload bench.py measures (BASELINE.json configs[1]-[4])
is rendered at its full size with its exact scene and switches, and row bands spread over the frame are compared
with the CPU checkers:

  * mixed100k (deterministic, 1 spp): the UNMODIFIED reference (oracle/_ref/ref_driver --rows --cols) when its
    prebuilt binary travelled, and the oracle port -- primary hit IDs bit-exact, 8-bit pixels within 1 LSB on
    >= 99.9 %, float image to rounding;
  * soup1m / glossy250k / dof4m (stochastic): the oracle port with the SAME Philox keys, so that every sample
    takes the same path -- hit IDs and per-class ray counts equal, pixels within 1 LSB on >= 99.9 %, float image
    to rounding (the reference itself draws from one serial mt19937 and cannot be compared sample by sample;
    tests/test_gpu_parity.py holds the stochastic effects to its 4096-spp renders);

once rendered by one rank and once re-assembled from the 8 interleaved-tile shards of a world of 8.
What must match: BVH::get_intersection / intersect_helper (acceleration.cpp:67-150), Trace + shade
(raytracer.cpp:180-351), the frame loop (raytracer.cpp:433-476).
"""
import numpy as np
import pytest

from conftest import GOLDEN, linear_mismatch, lsb_agreement

pytestmark = pytest.mark.gpu

# windows = (x0, y0, x1, y1) as fractions of the frame; sized so that the oracle needs seconds
BANDS = {
    "mixed100k": dict(rows=16, cols=1.0, at=(0.42, 0.60, 0.80)),
    "soup1m": dict(rows=4, cols=0.40, at=(0.35, 0.55, 0.75)),
    "glossy250k": dict(rows=4, cols=0.25, at=(0.35, 0.60, 0.85)),
    "dof4m": dict(rows=4, cols=0.10, at=(0.40, 0.60, 0.85)),
}


def windows(name, width, height):
    b = BANDS[name]
    out = []
    for i, f in enumerate(b["at"]):
        y0 = int(f * height)
        w = max(8, int(b["cols"] * width))
        x0 = 0 if w >= width else int((0.1 + 0.2 * i) * width)
        out.append((x0, y0, min(width, x0 + w), y0 + b["rows"]))
    return out


@pytest.fixture(scope="module")
def loaded(rt, oracle_mod):
    """name -> (scene dict, json path, product scene, oracle scene), built once per workload and dropped after its tests."""
    from oracle import scene_io
    from ray_tracying_b200 import workloads
    cache = {}

    def get(name):
        if name not in cache:
            cache.clear()  # one big scene at a time
            sc = workloads.scene_dict(name)
            path = workloads.scene_path_for(name, sc)
            scene = rt.Scene.from_json(path, GOLDEN)  # the product's own scene.json loader + BVH build, as in bench.py
            oracle = oracle_mod.OracleScene(*scene_io.scene_arrays(sc, GOLDEN))
            cache[name] = (sc, path, scene, oracle)
        return cache[name]

    yield get
    cache.clear()


@pytest.mark.parametrize("name", ["mixed100k", "soup1m", "glossy250k", "dof4m"])
def test_benchmark_workload_matches_the_cpu_checkers_on_row_bands(rt, oracle_mod, loaded, name):
    from ray_tracying_b200 import dist, workloads
    sc, path, scene, oracle = loaded(name)
    R = workloads.WORKLOADS[name]["render"]
    spp = max(1, R["n_samples_sqrt"]) ** 2 if R["n_samples_sqrt"] > 1 else 1
    width, height = scene.resolution
    assert (width, height) == tuple(workloads.WORKLOADS[name]["gen"][1]["resolution"])
    tile = (32, 32)

    # the frame exactly as bench.py renders it (seed 1, random shutter time), by one rank ...
    rgb, ids, lin, st = scene.render(seed=1, want_linear=True, tile=tile, **R)
    assert st.primary_rays == width * height * spp and (ids >= 0).mean() > 0.3
    # ... and re-assembled from the shards of a world of 8
    owner = dist.tile_owner(width, height, tile, 8)
    rgb8, ids8, lin8, rays8 = np.zeros_like(rgb), np.full_like(ids, -7), np.zeros_like(lin), [0, 0, 0]
    for r in range(8):
        part = scene.render(seed=1, want_linear=True, tile=tile, rank=r, world=8, **R)
        m = owner == r
        rgb8[m], ids8[m], lin8[m] = part[0][m], part[1][m], part[2][m]
        for k, v in enumerate((part[3].primary_rays, part[3].shadow_rays, part[3].secondary_rays)):
            rays8[k] += v
    assert np.array_equal(ids8, ids) and np.array_equal(rgb8, rgb) and np.array_equal(lin8.view(np.uint32), lin.view(np.uint32))
    assert tuple(rays8) == (st.primary_rays, st.shadow_rays, st.secondary_rays)

    hit_pixels = 0
    for win in windows(name, width, height):
        x0, y0, x1, y1 = win
        sl = (slice(y0, y1), slice(x0, x1))
        ref = oracle.render(seed=1, rows=(y0, y1), cols=(x0, x1), **R)
        assert np.array_equal(ids[sl], ref["ids"][sl]), f"{name} {win}: {int((ids[sl] != ref['ids'][sl]).sum())} primary hit IDs differ from the oracle"
        assert lsb_agreement(rgb[sl], ref["rgb"][sl], 1) >= 0.999, f"{name} {win}"
        worst = linear_mismatch(lin[sl], ref["linear"][sl], spp)
        assert worst <= 1.0, f"{name} {win}: float image off by {worst:.2f}x the rounding bound"
        # the same window rendered on its own: identical pixels, and the ray counts of exactly these pixels
        part = scene.render(seed=1, want_linear=True, tile=tile, window=win, **R)
        assert np.array_equal(part[2][sl].view(np.uint32), lin[sl].view(np.uint32)) and np.array_equal(part[1][sl], ids[sl])
        assert (part[3].primary_rays, part[3].shadow_rays, part[3].secondary_rays) == ref["rays"], f"{name} {win}: ray counts differ"
        hit_pixels += int((ref["ids"][sl] >= 0).sum())
        if name == "mixed100k" and oracle_mod.RefDriver.available():
            # the unmodified reference binary on the same band (deterministic workload)
            ids_ref, _, _ = oracle_mod.RefDriver.ids(path, use_bvh=True, rows=(y0, y1), cols=(x0, x1))
            rgb_ref, lin_ref, _ = oracle_mod.RefDriver.render(path, rows=(y0, y1), cols=(x0, x1), seed=1, **R)
            assert np.array_equal(ids[sl], ids_ref[:, x0:x1]), f"{name} {win}: hit IDs differ from the reference binary"
            assert lsb_agreement(rgb[sl], rgb_ref[:, x0:x1], 1) >= 0.999
            assert linear_mismatch(lin[sl], lin_ref[:, x0:x1], spp) <= 1.0
    assert hit_pixels > 0, "the sampled bands see no geometry: move them"
