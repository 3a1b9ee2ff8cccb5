"""First GPU sanity run: GPU render of the ASCII scene vs the reference driver on the same box."""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracying_b200 as rt

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene_path = os.path.join(root, "tests/golden/ascii_scene.json")
out = os.path.join(root, "gpurun_out"); os.makedirs(out, exist_ok=True)
ref = os.path.join(root, "oracle/_ref/ref_driver")

s = rt.Scene.from_json(scene_path, "/nonexistent")
print("devices", rt.device_count(), s.counts())
for prune in (True, False):
    rgb, ids, lin, st = s.render(use_bvh=True, n_samples_sqrt=1, prune=prune, want_linear=True)
    print("prune", prune, st)
    np.save(os.path.join(out, f"gpu_ids_prune{int(prune)}.npy"), ids)
rgb_lin, ids_lin, _, st = s.render(use_bvh=False, n_samples_sqrt=1)
print("linear", st)
subprocess.run([ref, "--scene", scene_path, "--mode", "ids", "--bvh", "1", "--out-ids", out + "/ref_ids.bin"], check=True, stderr=subprocess.DEVNULL)
raw = np.fromfile(out + "/ref_ids.bin", dtype=np.int32)
w, h = raw[0], raw[1]
ref_ids = raw[4:4 + w * h].reshape(h, w)
print("ids mismatches (bvh, prune):", int((ids != ref_ids).sum()), "of", ids.size)
ids0 = np.load(os.path.join(out, "gpu_ids_prune0.npy"))
print("ids mismatches (bvh, no prune):", int((ids0 != ref_ids).sum()))
subprocess.run([ref, "--scene", scene_path, "--mode", "ids", "--bvh", "0", "--out-ids", out + "/ref_ids_lin.bin"], check=True, stderr=subprocess.DEVNULL)
raw = np.fromfile(out + "/ref_ids_lin.bin", dtype=np.int32)
ref_ids_lin = raw[4:4 + w * h].reshape(h, w)
print("ids mismatches (linear):", int((ids_lin != ref_ids_lin).sum()))
print("ref bvh vs ref linear differ:", int((ref_ids != ref_ids_lin).sum()))
rt.write_ppm(out + "/gpu_ascii_1spp.ppm", rgb)
# timing of a few configs
for spp, ls in ((1, 1), (2, 1), (4, 4)):
    t0 = time.time(); _, _, _, st = s.render(use_bvh=True, n_samples_sqrt=spp, light_samples=ls, want_ids=False); t1 = time.time()
    print(f"s={spp} ls={ls}: kernel {st.kernel_ms:.2f} ms, rays {st.rays}, {st.rays / st.kernel_ms * 1e-3:.1f} Mrays/s, wall {t1 - t0:.3f}s")
