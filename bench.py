#!/usr/bin/env python
"""bench.py -- throughput of the render hot path (Mrays/s, ms/frame) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one frame of the workload: per-pixel ray generation, BVH traversal, ray-primitive
intersection, Blinn-Phong shading with shadow/reflection/refraction rays (the reference's frame
loop, raytracer.cpp:433-476). With N > 1 (torchrun, one rank per GPU) the frame is sharded by
interleaved screen tiles, scene and BVH replicated, and assembled with one all_gather at frame
end. Rays = get_intersection calls (primary + shadow + reflection + refraction).

Workloads (BASELINE.json configs; the default is configs[1]):
    mixed100k : 100k-shape mixed scene (spheres/ellipsoids, cubes incl. rod-like ones, rectangles,
                plane quads = 2 triangles each), 1920x1080, 1 spp, Whitted depth 5       [configs[1]]
    soup1m    : 1M-triangle soup (500k plane quads), 1080p, 64 spp, 16-sample area light  [configs[2]]
    glossy250k: 250k-triangle glossy scene, 3840x2160, 100 spp, depth 8                   [configs[3]]
    dof4m     : 4M-triangle scene, thin lens + motion blur, 3840x2160, 256 spp            [configs[4]]
    ascii     : the reference's own ASCII/scene.json, 1 spp                               [configs[0]]

--impl reference times the reference's OWN CPU code (oracle/_ref/ref_driver = the unmodified
reference sources behind a small driver) on a bounded sample of the same workload with every host
core (one process per core over row bands: the reference BVH object is not thread-safe).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CACHE = os.environ.get("RT_BENCH_CACHE", "/tmp/rt_b200_bench")

WORKLOADS = {
    # name: (generator kwargs, render kwargs, description)
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


def scene_path_for(name: str) -> str:
    os.makedirs(CACHE, exist_ok=True)
    if name == "ascii":
        return os.path.join(ROOT, "tests", "golden", "ascii_scene.json")
    path = os.path.join(CACHE, name + ".json")
    if not os.path.exists(path):
        from ray_tracying_b200 import scenes
        fn, kw = WORKLOADS[name]["gen"]
        t0 = time.time()
        sc = getattr(scenes, fn)(**kw)
        scenes.write_scene(sc, path + ".tmp")
        os.replace(path + ".tmp", path)
        print(f"[bench] generated {name}: {scenes.shape_count(sc)} shapes, {os.path.getsize(path) / 1e6:.1f} MB in {time.time() - t0:.1f}s",
              file=sys.stderr)
    return path


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=10)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref/ref_driver; oracle port only to COUNT rays)
# ------------------------------------------------------------------------------------------------
def reference_sample(scene_path: str, render: dict, repeats: int, target_seconds: float = 4.0):
    """Runs the unmodified reference on row bands spread over the frame, one process per host core.
    Returns dict(seconds=[per repeat, max over processes], rays, rows, cores, kind)."""
    from oracle.oracle import REF_DRIVER, OracleScene, RefDriver
    with open(scene_path) as f:
        scene = json.load(f)
    width, height = scene["render"]["resolution_x"], scene["render"]["resolution_y"]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, height))
    spp = max(1, render["n_samples_sqrt"]) ** 2 if render["n_samples_sqrt"] > 1 else 1
    kind = "reference" if RefDriver.available() else "port"

    def bands(rows_per_proc):
        stride = height / procs
        out = []
        for i in range(procs):
            y0 = min(height - rows_per_proc, int(i * stride + 0.5 * max(0.0, stride - rows_per_proc)))
            out.append((y0, y0 + rows_per_proc))
        return out

    def run_ref(rows_list, reps):
        cmds = []
        for (y0, y1) in rows_list:
            cmds.append([REF_DRIVER, "--scene", scene_path, "--mode", "render", "--bvh", str(int(render["use_bvh"])),
                         "--s", str(render["n_samples_sqrt"]), "--light-samples", str(render["light_samples"]),
                         "--depth", str(render["max_depth"]), "--seed", "1", "--rows", str(y0), str(y1), "--repeat", str(reps)])
        ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for c in cmds]
        per_proc = []
        for p in ps:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError("ref_driver failed")
            info = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
            per_proc.append(info["all_seconds"])
        return [max(pp[i] for pp in per_proc) for i in range(reps)], per_proc

    oracle = OracleScene.from_dict(scene, os.path.join(ROOT, "tests", "golden"))

    def run_port(rows_list, reps):
        secs = []
        for _ in range(reps):
            t0 = time.perf_counter()
            for rows in rows_list:
                oracle.render(rows=rows, threads=cores, seed=1, **render)
            secs.append(time.perf_counter() - t0)
        return secs, None

    run = run_ref if kind == "reference" else run_port
    # calibrate: one row per process, then scale the band to ~target_seconds per repeat
    cal, _ = run(bands(1), 1)
    rows_per_proc = int(max(1, min(height // procs, round(target_seconds / max(cal[0], 1e-3)))))
    rows_list = bands(rows_per_proc)
    seconds, _ = run(rows_list, repeats)
    # ray count of exactly these rows from the oracle port (identical for deterministic workloads,
    # statistically equal for stochastic ones)
    rays = 0
    for rows in rows_list:
        rays += sum(oracle.render(rows=rows, threads=cores, seed=1, **render)["rays"])
    return {"seconds": seconds, "rays": rays, "rows": rows_per_proc * procs, "cores": procs if kind == "reference" else cores,
            "kind": kind, "width": width, "height": height, "spp": spp}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    path = scene_path_for(args.workload)
    # bounded: the whole run (warm-up + steps samples, each on every host core) stays within ~2 minutes
    res = reference_sample(path, wl["render"], args.warmup + args.steps, target_seconds=max(0.5, min(4.0, 100.0 / (args.warmup + args.steps))))
    secs = res["seconds"][args.warmup:]
    ms = float(np.mean(secs)) * 1e3
    value = res["rays"] / (ms * 1e-3) / 1e6
    sample = (f"{res['rows']} of {res['height']} rows x {res['width']} px x {res['spp']} spp per step, "
              f"row bands spread over the frame, one process per core")
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, **wl["render"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_step": res["rays"],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import ray_tracying_b200 as rt
    from ray_tracying_b200 import dist as rdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    if rt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        path = scene_path_for(args.workload)
    barrier()
    path = scene_path_for(args.workload)

    t0 = time.perf_counter()
    scene = rt.Scene.from_json(path, os.path.join(ROOT, "tests", "golden"))
    load_s = time.perf_counter() - t0
    width, height = scene.resolution
    tile = (32, 32)
    R = wl["render"]
    params = rt.make_params(rank=rank, world=world, tile=tile, seed=1, **R)
    h2d_bytes = scene.upload()

    rgb = torch.zeros((height, width, 3), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    # one untimed pass with counters: rays per step and the traversal work (for the roofline)
    st = scene.render_device(rt.make_params(rank=rank, world=world, tile=tile, seed=1, collect_stats=True, **R), rgb.data_ptr(),
                             0, 0, stream.cuda_stream)
    counts = torch.tensor([st.rays, st.primary_rays, st.shadow_rays, st.secondary_rays, st.node_visits, st.prim_tests],
                          dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    rays, n_primary, n_shadow, n_secondary, node_visits, prim_tests = (int(x) for x in counts.tolist())
    launches_per_step = int(st.launches)

    params = rt.make_params(rank=rank, world=world, tile=tile, seed=1, **R)
    # roofline leg: the same frames with the launches serialised on one stream and CUDA events
    # around every trace / shadow / shade / light launch, so each kernel's duration is its own
    # (in the headline steps shadow/light of a level overlap trace/shade of the next level)
    params_serial = rt.make_params(rank=rank, world=world, tile=tile, seed=1, time_kernels=True, serial=True, **R)

    def one_step():
        scene.render_device(params, rgb.data_ptr(), 0, 0, stream.cuda_stream, sync_stats=False)
        return rdist.gather_frame(rgb, width, height, tile, rank, world) if world > 1 else rgb

    clocks = ClockSampler(local_rank)
    step_ms, kernel_ms = [], []
    for i in range(args.warmup + args.steps):
        flush.zero_()  # evict the scene from L2 between iterations
        if i == args.warmup and rank == 0:
            clocks.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), scene.last_timing()[0]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= args.warmup:
            step_ms.append(float(t[0]))
            kernel_ms.append(float(t[1]))
    trav_ms, serial_ms, trav_launches, class_ms = [], [], 0, {}
    for i in range(max(3, min(args.steps, 10))):
        flush.zero_()
        barrier()
        scene.render_device(params_serial, rgb.data_ptr(), 0, 0, stream.cuda_stream, sync_stats=False)
        barrier()
        kt = scene.last_kernel_times()
        t = torch.tensor([kt["trace"][0] + kt["shadow"][0], scene.last_timing()[0]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        trav_ms.append(float(t[0]))
        serial_ms.append(float(t[1]))
        trav_launches = kt["trace"][1] + kt["shadow"][1]
        trav_frame_launches = kt["trace"][2] + kt["shadow"][2]
        for k, (ms_k, n_k, tot_k) in kt.items():
            class_ms.setdefault(k, []).append(ms_k * tot_k / max(n_k, 1))  # scaled to the whole frame
    clock_info = clocks.stop() if rank == 0 else {}
    ms = float(np.mean(step_ms))
    k_ms = float(np.mean(kernel_ms))
    value = rays / (ms * 1e-3) / 1e6

    # end to end through the C ABI with HOST buffers, every step: H2D copy of the scene from
    # page-locked host memory, render, D2H copy of the frame into page-locked host memory.
    #   1 GPU : rt_render() does all of it (Scene.render_into);
    #   N GPUs: upload + rt_render_device + frame-end all_gather + D2H of the assembled frame.
    host_frame = torch.empty((height, width, 3), dtype=torch.uint8, pin_memory=True)
    p_e2e = rt.make_params(rank=rank, world=world, tile=tile, seed=1, **R)
    e2e_s = []
    e2e_warm = max(1, min(args.warmup, 2))
    for i in range(e2e_warm + args.steps):
        flush.zero_()
        scene.evict()
        barrier()
        w0 = time.perf_counter()
        if world == 1:
            scene.render_into(p_e2e, host_frame.data_ptr())
        else:
            scene.upload()
            scene.render_device(p_e2e, rgb.data_ptr(), 0, 0, stream.cuda_stream, sync_stats=False)
            frame = rdist.gather_frame(rgb, width, height, tile, rank, world)
            host_frame.copy_(frame, non_blocking=False)
            torch.cuda.synchronize()
        w1 = time.perf_counter()
        t = torch.tensor([w1 - w0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= e2e_warm:
            e2e_s.append(float(t[0]))
    e2e_value = rays / float(np.mean(e2e_s)) / 1e6

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        # Dominant kernel = the traversal loop (trace_kernel and shadow_kernel are the two
        # instantiations of wave_loop). Algorithmic bytes: every box test reads
        # one child box (32 B of a 128 B node), every primitive test the 64 B head of a primitive
        # record, every ray its 32 B origin/direction and writes a 4 B result.
        alg_bytes = node_visits * 32 + prim_tests * 64 + rays * 36
        n_launch = max(1, trav_frame_launches)
        # frames with many batches time the first 512 launches only: scale to the frame's launch count
        trav = float(np.mean(trav_ms)) * n_launch / max(1, trav_launches)
        achieved = alg_bytes / world / (trav * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        # what actually bounds the loop (committed ncu capture of this workload, not measured in this run):
        # issue-slot utilisation and active lanes per instruction of the traversal kernels
        ncu_note = None
        kp = os.path.join(ROOT, "profiles", "ncu_key_metrics.json")
        if os.path.exists(kp) and args.workload == "mixed100k":
            with open(kp) as f:
                km = json.load(f)
            ncu_note = {"source": km["source"],
                        "kernels": {k.replace("void ", ""): {m: round(v[0][m], 2) for m in ("issue_slot_pct", "lanes_per_instruction", "l1_hit_pct", "warps_active_pct")}
                                    for k, v in km["kernels"].items()}}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, **R, "parallelism": f"tiles{world}", "tile": list(tile),
                       "l2": "flushed between steps (256 MiB memset)", "shapes": scene.counts()["shapes"],
                       "resolution": [width, height]},
            "rays_per_step": rays, "rays": {"primary": n_primary, "shadow": n_shadow, "secondary": n_secondary},
            "kernel_ms_per_step": k_ms, "kernel_mrays_per_s": rays / (k_ms * 1e-3) / 1e6,
            "kernel_class_ms_per_step": {k: float(np.mean(v)) for k, v in class_ms.items()},
            "gpu_launches": launches_per_step * args.steps,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(width * height * 3), "ms_per_step": float(np.mean(e2e_s)) * 1e3,
                    "path": "rt_render (C ABI, pinned host buffers)" if world == 1 else
                            "rt_scene_upload + rt_render_device + all_gather + D2H"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "wave_loop = trace_kernel + shadow_kernel (two instantiations of one traversal loop)",
                         "launches_per_step": n_launch, "timed_launches_per_step": trav_launches, "avg_launch_ms": trav / n_launch,
                         "algorithmic_bytes_per_launch": alg_bytes // world // n_launch,
                         "share_of_step": trav / float(np.mean(serial_ms)), "serialised_step_ms": float(np.mean(serial_ms)),
                         "timing": "CUDA events around each launch, launches serialised on one stream (roofline leg)",
                         "note": "scene is L2/L1 resident: DRAM traffic is a few % of the algorithmic bytes, the loop is issue-bound",
                         "ncu": ncu_note,
                         "box_tests_per_ray": node_visits / max(rays, 1), "prim_tests_per_ray": prim_tests / max(rays, 1)},
            "clocks": clock_info,
            "host": {"scene_load_and_bvh_build_s": load_s},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                res = reference_sample(path, R, repeats=2, target_seconds=3.0)
                cms = float(np.mean(res["seconds"][1:])) * 1e3
                line["cpu_baseline"] = {
                    "value": res["rays"] / (cms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"],
                    "sample": f"{res['rows']} of {res['height']} rows x {res['width']} px x {res['spp']} spp, row bands spread over the frame"}
            except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="mixed100k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
