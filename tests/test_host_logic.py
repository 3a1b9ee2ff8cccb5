"""CPU tests of the host side of the drop-in boundary: the C ABI loads and exports every symbol
include/rt_render.h declares, the C++ scene loader agrees with an independent Python restatement
of the reference loader, the BVH equals the reference's tree, PPM I/O, argument checking, and the
'no GPU -> fail loudly' rule. No compute calls are made here."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_data, golden_scene, scene_file, with_resolution


def test_library_exports_every_declared_symbol(rt):
    header = open(os.path.join(ROOT, "include", "rt_render.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = C.CDLL(os.path.join(ROOT, "ray_tracying_b200", "librt_b200.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"librt_b200.so does not export {name}"
    from ray_tracying_b200 import _lib
    assert declared == set(_lib.SYMBOLS), "python binding and header disagree"


def test_struct_sizes_match_header(rt):
    from ray_tracying_b200 import _lib
    assert C.sizeof(_lib.CameraDesc) == 4 * (9 + 1 + 2 + 2 + 2)
    assert C.sizeof(_lib.LightDesc) == 32
    assert C.sizeof(_lib.MaterialDesc) == 60
    assert C.sizeof(_lib.ShapeDesc) == 104
    assert C.sizeof(_lib.RenderParams) == 80
    assert C.sizeof(_lib.RenderStats) == 64


def test_defaults_are_the_reference_cli_defaults(rt):
    p = rt.make_params()
    # raytracer.cpp:361-363 and raytracer.hpp:11
    assert (p.use_bvh, p.samples_sqrt, p.light_samples, p.max_depth) == (0, 4, 1, 10)
    assert (p.rank, p.world) == (0, 1)


@pytest.mark.parametrize("name", ["mixed_400", "few_3", "few_5", "ties_axis_aligned", "numerics_edge", "textured_40", "empty"])
def test_bvh_equals_reference_tree(rt, name):
    """Tree topology, leaf contents and every box bit-equal to the reference's (golden dump)."""
    scene = rt.Scene.from_json(os.path.join(GOLDEN, name + ".json"), GOLDEN)
    ref = json.loads(str(golden_data(name)["bvh_dump"]))
    mine = scene.dump_bvh()
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a[0] == b[0] and a[3] == b[3]
        assert np.array_equal(np.float32(a[1]), np.float32(b[1])) and np.array_equal(np.float32(a[2]), np.float32(b[2]))


def test_ascii_scene_counts_and_tree(rt, tmp_path):
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "ascii_scene.json"), GOLDEN)
    assert scene.resolution == (1920, 1080)
    c = scene.counts()
    assert c["shapes"] == 141 and c["lights"] == 2 and c["materials"] == 1
    ref = json.loads(str(golden_data("ascii_scene_480")["bvh_dump"]))
    assert [(n[0], n[3]) for n in scene.dump_bvh()] == [(n[0], n[3]) for n in ref]


def test_loader_matches_python_restatement(rt, oracle_mod, tmp_path):
    """C++ JSON path (rt_scene_load_json) and the array path fed by oracle/scene_io.py build the same scene."""
    from oracle import scene_io
    d = golden_scene("textured_40")
    cam, lights, mats, shapes, tex = scene_io.scene_arrays(d, GOLDEN)
    a = rt.Scene.from_json(os.path.join(GOLDEN, "textured_40.json"), GOLDEN)
    mats15 = np.concatenate([mats["f"], mats["texture"].astype(np.float32)[:, None]], axis=1)
    b = rt.Scene.from_arrays(cam, lights.view(np.float32).reshape(-1, 8), mats15, shapes.view(rt.SHAPE_DTYPE), tex)
    assert a.counts() == b.counts()
    assert np.array_equal(a.shape_order(), b.shape_order())
    assert a.dump_bvh() == b.dump_bvh()
    o = oracle_mod.OracleScene(cam, lights, mats, shapes, tex)
    assert np.array_equal(o.shape_order(), a.shape_order())


def test_sensor_size_is_truncated_like_the_reference(rt, tmp_path):
    # camera.cpp:39-40 reads sensor_width/height with get<int>()
    d = golden_scene("few_3")
    d["cameras"][0]["sensor_width"] = 36.9
    p = scene_file(tmp_path, d)
    from oracle import scene_io
    cam = scene_io.load_scene(p, GOLDEN)[0]
    assert cam["sensor_width"] == 36
    rt.Scene.from_json(p, GOLDEN)


def test_invalid_entries_are_skipped_like_the_reference(rt, tmp_path):
    d = golden_scene("few_3")
    n0 = rt.Scene.from_json(scene_file(tmp_path, d), GOLDEN).counts()
    d["cubes"] = d.get("cubes", []) + [{"rotation": [0, 0, 0]}, 7]           # no translation / not an object
    d["planes"] = d.get("planes", []) + [{"corners": [[0, 0, 0]]}]            # not 4 corners
    d["lights"] = d["lights"] + [{"location": [0, 0, 1], "color": [1, 1, 1], "intensity": 0.0}, {"color": [1, 1, 1]}]
    n1 = rt.Scene.from_json(scene_file(tmp_path, d, "b.json"), GOLDEN).counts()
    assert n1["shapes"] == n0["shapes"] and n1["lights"] == n0["lights"]


def test_missing_file_and_bad_json_report_errors(rt, tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_json(str(tmp_path / "nope.json"))
    assert e.value.status == -2
    bad = tmp_path / "bad.json"
    bad.write_text("{\"cameras\": [")
    with pytest.raises(rt.RtError):
        rt.Scene.from_json(str(bad))
    nocam = tmp_path / "nocam.json"
    nocam.write_text("{\"lights\": []}")
    with pytest.raises(rt.RtError):
        rt.Scene.from_json(str(nocam))


def test_parallel_loader_builds_the_scene_a_serial_loader_builds(rt, tmp_path):
    """The four shape arrays are parsed and converted element by element on several threads; materials, textures and
    the primitives' order are settled serially in element order. 30k shapes (enough for several chunks) with
    invalid entries sprinkled in: the scene (shape order after BVH construction, tree, counts) is identical with one
    host thread and with many, and identical to the oracle's independent Python loader + C++ builder."""
    from ray_tracying_b200 import scenes
    sc = scenes.mixed_scene(30000, seed=5, resolution=(64, 36), texture_file="checker.jpg")
    # strings with brackets, quotes and backslashes inside the big arrays: the multi-threaded bracket scan must see through them
    weird = dict(sc["spheres"][7]["material"], texture_file='a]b[{"x\\.jpg')
    for i in (3, 2000, 7000, 10400):
        sc["spheres"][i] = dict(sc["spheres"][i], material=weird, note='}],[{ \\" ]')
    sc["cubes"][5000] = dict(sc["cubes"][5000], note='\\\\"]')
    sc["planes"][100:100] = [{"corners": [[0, 0, 0]]}, 7, {"corners": "x"}]
    sc["spheres"][50:50] = [{"rotation": [0, 0, 0]}, {"location": [0, 0, 0], "material": {"diffuse_color": [1, 2]}}]
    sc["cubes"][10:10] = [{"rotation": [0, 0, 0]}]
    path = scene_file(tmp_path, sc)
    code = ("import sys, json; sys.path.insert(0, %r)\n"
            "import numpy as np, ray_tracying_b200 as rt\n"
            "s = rt.Scene.from_json(%r, %r)\n"
            "np.save(%r + sys.argv[1] + '.npy', s.shape_order()); print(json.dumps(s.counts()))\n") % (ROOT, path, GOLDEN, str(tmp_path / "order_"))
    outs = {}
    for threads in ("1", "7"):
        r = subprocess.run([os.sys.executable, "-c", code, threads], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           env=dict(os.environ, RT_B200_HOST_THREADS=threads), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[threads] = (json.loads(r.stdout.strip().splitlines()[-1]), np.load(str(tmp_path / ("order_" + threads + ".npy"))), r.stderr)
    assert outs["1"][0] == outs["7"][0] and np.array_equal(outs["1"][1], outs["7"][1])
    assert outs["1"][2] == outs["7"][2], "warnings come out in element order whatever the thread count"
    assert outs["1"][0]["shapes"] == 30001 and "Skipping invalid plane definition" in outs["1"][2]  # the sphere with a bad material keeps the default material
    from oracle import oracle, scene_io
    o = oracle.OracleScene(*scene_io.scene_arrays(sc, GOLDEN))
    assert np.array_equal(o.shape_order(), outs["7"][1])


def test_malformed_element_rejects_the_whole_document(rt, tmp_path):
    """A syntax error inside a shape array (whose elements are parsed on their own, later) still fails the load,
    like one parse of the whole document would; a duplicate key keeps its LAST value, arrays included."""
    d = golden_scene("few_3")
    text = json.dumps(d)
    key = '"planes": [' if '"planes": [' in text else '"spheres": ['
    bad = tmp_path / "bad_elem.json"
    bad.write_text(text.replace(key, key + '{"corners": [[0,0,0],[1,0,0],[1,1,0],[0,1,x]]}, ', 1))
    with pytest.raises(rt.RtError) as e:
        rt.Scene.from_json(str(bad), GOLDEN)
    assert e.value.status == -2
    n0 = rt.Scene.from_json(scene_file(tmp_path, d), GOLDEN).counts()["shapes"]
    dup = tmp_path / "dup.json"
    dup.write_text(text[:-1] + ', "spheres": [], "cubes": 3}')
    n1 = rt.Scene.from_json(str(dup), GOLDEN).counts()["shapes"]
    assert n1 == n0 - len(d.get("spheres", [])) - len(d.get("cubes", []))


def test_ppm_roundtrip_and_reference_format(rt, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    p = str(tmp_path / "a.ppm")
    rt.write_ppm(p, img)
    text = open(p).read().splitlines()
    assert text[0] == "P3" and text[1] == "7 5" and text[2] == "255"
    # Image::write separates pixels by two spaces and channels by one (image.cpp:66-78)
    assert text[3] == "  ".join(" ".join(str(int(v)) for v in px) for px in img[0])
    assert np.array_equal(rt.read_ppm(p), img)
    tex = rt.read_ppm(os.path.join(GOLDEN, "checker.ppm"))
    assert tex.shape == (32, 32, 3)


def test_shard_pixels_partition_the_frame(rt):
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "mixed_400.json"), GOLDEN)
    w, h = scene.resolution
    from ray_tracying_b200 import dist
    for world in (1, 2, 3, 8):
        for tile in ((32, 32), (64, 16), (8, 4)):
            counts = [scene.shard_pixels(rt.make_params(rank=r, world=world, tile=tile)) for r in range(world)]
            assert sum(counts) == w * h
            owner = dist.tile_owner(w, h, tile, world)
            assert counts == [int((owner == r).sum()) for r in range(world)]


def test_shard_pixels_with_tile_blocks_and_a_window(rt):
    """Tile blocks (B x B groups of tiles per owner) and the render window: the ranks still partition exactly the
    pixels that are rendered, and rt_shard_pixels agrees with dist.tile_owner."""
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "mixed_400.json"), GOLDEN)
    w, h = scene.resolution
    from ray_tracying_b200 import dist
    for world, tile, block in ((4, (8, 4), 3), (3, (16, 8), 2), (8, (32, 32), 4)):
        owner = dist.tile_owner(w, h, tile, world, block)
        counts = [scene.shard_pixels(rt.make_params(rank=r, world=world, tile=tile, tile_block=block)) for r in range(world)]
        assert counts == [int((owner == r).sum()) for r in range(world)] and sum(counts) == w * h
        win = (13, 7, w - 21, h - 30)
        inside = np.zeros((h, w), dtype=bool)
        inside[win[1]:win[3], win[0]:win[2]] = True
        counts = [scene.shard_pixels(rt.make_params(rank=r, world=world, tile=tile, tile_block=block, window=win)) for r in range(world)]
        assert counts == [int(((owner == r) & inside).sum()) for r in range(world)]
    with pytest.raises(rt.RtError):
        scene.shard_pixels(rt.make_params(window=(w + 5, 0, w + 9, 4)))  # empty after clipping to the frame
    with pytest.raises(ValueError):
        rt.make_params(window=(5, 5, 5, 9))


def test_bad_params_are_rejected(rt):
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "few_3.json"), GOLDEN)
    for kw in (dict(rank=2, world=2), dict(tile=(30, 32)), dict(tile=(32, 6))):
        with pytest.raises(rt.RtError) as e:
            scene.shard_pixels(rt.make_params(**kw))
        assert e.value.status == -1


def test_render_without_gpu_fails_loudly(rt):
    if rt.device_count() > 0:
        pytest.skip("a GPU is present")
    scene = rt.Scene.from_json(os.path.join(GOLDEN, "few_3.json"), GOLDEN)
    with pytest.raises(rt.RtError) as e:
        scene.render(use_bvh=True, n_samples_sqrt=1)
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(rt.RtError):
        scene.upload()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "ray_tracying_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "rt_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_cli_usage_error_matches_reference(rt):
    exe = os.path.join(ROOT, "ray_tracying_b200", "bin", "Raytracer")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1
    assert "Please specify scene file name" in r.stderr  # raytracer.cpp:392
    if rt.device_count() == 0:
        r = subprocess.run([exe, "-input", os.path.join(GOLDEN, "few_3.json"), "-output", "/tmp/x.ppm", "-bvh", "-s", "1"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.parametrize("name", ["mixed_400", "few_3", "few_5", "ties_axis_aligned", "numerics_edge", "textured_40", "ascii_scene"])
def test_wide_tree_covers_every_primitive_once(rt, name):
    """The flattened 4-wide device tree (csrc/scene.hpp DWide) is a BVH over the primitives' culling boxes: every
    sorted position is the child of exactly one node; the inner children of a node are consecutive nodes in the
    leading slots, primitive children follow; every child box contains what is below it -- the culling boxes, or,
    for spheres (whose culling test carries a distance-dependent term that must stay local), the box of the sphere's
    reference leaf, which every ray that may test the sphere passes; the culling boxes of a reference leaf's
    primitives cover that leaf's box (each contains its primitive's box, whose union the leaf box is); nodes are
    numbered breadth-first; no ray can need more than 36 stack entries."""
    scene = rt.Scene.from_json(os.path.join(GOLDEN, name + ".json"), GOLDEN)
    nodes, depth = scene.dump_wide()
    ref = scene.dump_bvh()  # pre-order reference nodes: (is_leaf, lo, hi, prims in load order)
    order = scene.shape_order()
    pos_of = {int(load): pos for pos, load in enumerate(order)}
    bits = nodes.view(np.uint32)
    seen, max_sp, levels = {}, [0], {}
    leaf_box_of = {}
    for is_leaf, lo, hi, prims in ref:
        if is_leaf:
            for p_ in prims:
                leaf_box_of[pos_of[p_]] = (np.float32(lo), np.float32(hi))

    def box(n, k):
        f = nodes[n]
        return np.array([f[0 + k], f[8 + k], f[16 + k]]), np.array([f[4 + k], f[12 + k], f[20 + k]])

    def visit(n, level, sp):
        """returns (lo, hi) of everything below node n; sp = stack entries in use when n is visited"""
        levels[n] = level
        first, meta = int(bits[n, 24]), int(bits[n, 25])
        valid, prim = meta & 15, (meta >> 4) & 15
        assert valid in (1, 3, 7, 15), "valid children occupy slots 0..n-1"
        cnt = bin(valid).count("1")
        ni = cnt - bin(prim).count("1")
        assert prim & ~valid == 0 and prim == (valid & ~((1 << ni) - 1)), "inner children lead, primitives follow"
        assert level <= depth
        max_sp[0] = max(max_sp[0], sp + max(0, ni - 1))
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        for k in range(cnt):
            lo, hi = box(n, k)
            assert (lo <= hi).all()
            if k < ni:
                clo, chi = visit(first + k, level + 1, sp + (ni - 1 - k if k < ni - 1 else 0))
                assert (lo <= clo).all() and (hi >= chi).all(), "a child box contains what is below it"
            else:
                pos = int(bits[n, 27 + k])
                assert pos not in seen, "a primitive appears once"
                seen[pos] = (lo, hi)
                if ((meta >> (16 + 2 * k)) & 3) == 0 and nodes[n, 26] > 0:  # sphere: the ancestors hold its leaf box
                    lo, hi = leaf_box_of[pos]
            lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
        return lo_all, hi_all

    if len(ref) == 0:
        assert len(nodes) == 0
        return
    visit(0, 1, 0)
    assert sorted(seen) == list(range(len(order))), "every sorted position is reachable exactly once"
    assert max_sp[0] + 1 <= 36
    bfs = [levels[n] for n in range(len(nodes))]
    assert bfs == sorted(bfs), "breadth-first numbering: levels do not decrease with the node index"
    for is_leaf, lo, hi, prims in ref:
        if not is_leaf:
            continue
        pos = [pos_of[p] for p in prims]
        assert pos == list(range(pos[0], pos[0] + len(pos))), "a reference leaf holds consecutive sorted positions"
        clo = np.min([seen[q][0] for q in pos], axis=0)
        chi = np.max([seen[q][1] for q in pos], axis=0)
        assert (clo <= np.float32(lo)).all() and (chi >= np.float32(hi)).all(), "culling boxes cover the leaf box"


def test_json_numbers_are_strtod_exact(rt):
    """The loader's number fast path (significand < 2^53 times / over a power of ten <= 1e22) must return
    the double Python's float() (= correctly rounded strtod) returns, bit for bit; integers stay integers."""
    import ctypes as C
    import random
    from ray_tracying_b200._lib import lib
    rng = random.Random(7)
    cases = ["0", "-0", "0.0", "-0.0", "1", "-17", "123456789012345678", "1234567890123456789", "0.1", "0.30000000000000004",
             "1e22", "1e23", "1E-22", "1e-23", "9007199254740991.0", "9007199254740993.0", "4.35", "0.000001", "1.7976931348623157e308",
             "5e-324", "2.2250738585072011e-308", "123456789012345678901234567890.5", "0.1000000000000000055511151231257827",
             "3.14159265358979323846264338327950288", "-2.5e-5", "1e0", "12.0e+3", "8.0E3", "0.5000"]
    for _ in range(20000):
        digits = rng.randint(1, 22)
        m = str(rng.randint(0, 10 ** digits - 1))
        cut = rng.randint(0, len(m))
        txt = (m[:cut] or "0") + ("." + m[cut:] if cut < len(m) else "")
        if rng.random() < 0.3:
            txt += "e" + str(rng.randint(-30, 30))
        if rng.random() < 0.5:
            txt = "-" + txt
        if txt.lstrip("-").startswith("0") and len(txt.lstrip("-")) > 1 and txt.lstrip("-")[1] != "." and txt.lstrip("-")[1] not in "eE":
            txt = txt.replace("0", "1", 1)  # JSON forbids leading zeros
        cases.append(txt)
    for txt in cases:
        val, is_int = C.c_double(), C.c_int32()
        assert lib.rt_json_number(txt.encode(), C.byref(val), C.byref(is_int)) == 0, txt
        plain_int = txt.lstrip("-").isdigit() and len(txt) < 19
        assert bool(is_int.value) == plain_int, txt
        if plain_int:  # nlohmann: a plain integer is an int64 ("-0" is the integer 0)
            assert val.value == float(int(txt)), txt
        else:
            want = float(txt)
            assert np.float64(val.value).view(np.uint64) == np.float64(want).view(np.uint64), (txt, val.value, want)
