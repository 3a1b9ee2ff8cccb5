#!/usr/bin/env python
"""bench.py -- throughput of the render hot path (Mrays/s, ms/frame) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A step = one frame of the workload: per-pixel ray generation, BVH traversal, ray-primitive
intersection, Blinn-Phong shading with shadow/reflection/refraction rays (the reference's frame
loop, raytracer.cpp:433-476). Rays = get_intersection calls (primary + shadow + reflection +
refraction). The headline workload is BASELINE.json's configs[2] -- the one its target is stated on:
1M-triangle soup, 1920x1080, 64 spp antialiasing, 16-sample area-light soft shadows -- at every N; the
same JSON line carries configs[1] (100k mixed shapes, 1 spp, depth 5: a 5 ms frame) as `secondary`.

  value : device-timed. With N > 1 (torchrun, one rank per GPU) the frame is sharded by interleaved
          screen tiles, scene and BVH replicated, no traffic between GPUs inside the render loop, one
          NCCL all_gather of the packed tiles at frame end; max over ranks of the CUDA-event time.
  e2e   : the same frame through the C ABI with HOST buffers, every step: scene evicted, H2D copy of
          the scene from page-locked memory, render, D2H copy of the frame into page-locked memory.
          N = 1: rt_render. N > 1: rt_render_multi -- ONE process (rank 0) drives the N GPUs, each GPU
          copies its own packed tiles to the host, no NCCL; the other ranks wait.
  roofline : the traversal kernels against a MEASURED traversal ceiling of this chip (rt_traversal_peak:
          the loop's node step with fully converged warps on L1-resident nodes), in box tests per second.

--impl reference times the reference's OWN CPU code (oracle/_ref/ref_driver = the unmodified
reference sources behind a small driver) on a bounded sample of the same workload with every host
core (one process per core over windows spread over the frame: the reference BVH object is not
thread-safe). Nothing of the product library is loaded in that arm.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from ray_tracying_b200 import workloads  # noqa: E402  (pure Python: does not load librt_b200.so)

WORKLOADS = workloads.WORKLOADS
scene_path_for = workloads.scene_path_for


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
# the reference on the host cores (oracle/_ref/ref_driver; oracle port only to size the sample and COUNT rays)
# ------------------------------------------------------------------------------------------------
def reference_sample(name: str, scene_path: str, render: dict, repeats: int, target_seconds: float = 4.0):
    """Runs the unmodified reference on windows spread over the frame, one process per host core, ONE
    invocation per process (the scene load + BVH build of the reference takes ~40 s for the 1M-triangle
    scene). Returns dict(seconds=[per repeat, max over processes], rays, pixels, cores, kind, ...)."""
    from oracle import scene_io
    from oracle.oracle import REF_DRIVER, OracleScene, RefDriver
    sc = workloads.scene_dict(name)
    width, height = sc["render"]["resolution_x"], sc["render"]["resolution_y"]
    spp = max(1, render["n_samples_sqrt"]) ** 2 if render["n_samples_sqrt"] > 1 else 1
    kind = "reference" if RefDriver.available() else "port"
    oracle = OracleScene(*scene_io.scene_arrays(sc, os.path.join(ROOT, "tests", "golden")))
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, height))
    if kind == "reference":
        # nlohmann's DOM of the scene file costs ~12x its size per process: keep the processes within half the RAM
        try:
            import psutil
            procs = max(1, min(procs, int(0.5 * psutil.virtual_memory().available / max(1.0, 12.0 * os.path.getsize(scene_path)))))
        except Exception:
            pass

    # rays per pixel from a probe with the port, then a window per process worth ~target_seconds of one core
    # (0.07-0.11 Mrays/s per core is what the reference does on the GPU boxes' hosts; the time is MEASURED below)
    probe_rows = (int(0.55 * height), int(0.55 * height) + 1)
    probe_cols = (int(0.4 * width), int(0.4 * width) + 64)
    probe = oracle.render(rows=probe_rows, cols=probe_cols, seed=1, **render)
    rays_per_px = max(1.0, sum(probe["rays"]) / 64.0)
    target_px = max(8, int(0.08e6 * target_seconds * (1.0 if kind == "reference" else cores) / rays_per_px))
    rows = max(1, min(height // procs, target_px // width))
    cols = width if target_px >= width else max(8, target_px)
    wins = []
    for i in range(procs):
        stride = height / procs
        y0 = min(height - rows, int(i * stride + 0.5 * max(0.0, stride - rows)))
        x0 = 0 if cols >= width else int((i * 0.618034) % 1.0 * (width - cols))  # golden-ratio spread over the columns
        wins.append((x0, y0, x0 + cols, y0 + rows))

    if kind == "reference":
        cmds = [[REF_DRIVER, "--scene", scene_path, "--mode", "render", "--bvh", str(int(render["use_bvh"])),
                 "--s", str(render["n_samples_sqrt"]), "--light-samples", str(render["light_samples"]),
                 "--depth", str(render["max_depth"]), "--seed", "1", "--rows", str(y0), str(y1), "--cols", str(x0), str(x1),
                 "--repeat", str(repeats)] for (x0, y0, x1, y1) in wins]
        ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for c in cmds]
        per_proc = []
        for p in ps:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError("ref_driver failed")
            per_proc.append(json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])["all_seconds"])
        seconds = [max(pp[i] for pp in per_proc) for i in range(repeats)]
        used = procs
    else:
        seconds = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            for (x0, y0, x1, y1) in wins:
                oracle.render(rows=(y0, y1), cols=(x0, x1), threads=cores, seed=1, **render)
            seconds.append(time.perf_counter() - t0)
        used = cores
    # ray count of exactly these windows from the oracle port (identical for deterministic workloads,
    # statistically equal for stochastic ones: the reference draws from mt19937, the port from Philox)
    rays = sum(sum(oracle.render(rows=(y0, y1), cols=(x0, x1), threads=cores, seed=1, **render)["rays"]) for (x0, y0, x1, y1) in wins)
    return {"seconds": seconds, "rays": rays, "pixels": rows * cols * procs, "window": [cols, rows], "cores": used,
            "kind": kind, "width": width, "height": height, "spp": spp}


def sample_text(res: dict) -> str:
    return (f"{res['cores']} windows of {res['window'][0]} x {res['window'][1]} px spread over the {res['width']} x {res['height']} frame "
            f"({res['pixels']} px x {res['spp']} spp, {res['rays']} rays per step), one process per core")


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    path = scene_path_for(args.workload)
    # bounded: the whole run (warm-up + steps samples, each on every host core) stays within a few minutes
    res = reference_sample(args.workload, path, wl["render"], args.warmup + args.steps,
                           target_seconds=max(0.5, min(4.0, 100.0 / (args.warmup + args.steps))))
    secs = res["seconds"][args.warmup:]
    ms = float(np.mean(secs)) * 1e3
    value = res["rays"] / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, **wl["render"], "sample": sample_text(res)},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"], "sample": sample_text(res)},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_step": res["rays"],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def kernel_source_tag() -> str:
    """Identifies the kernel sources a profile under profiles/ was captured with."""
    h = hashlib.sha256()
    for f in ("render.cu", "rt_device.cuh", "philox.cuh", "bvh.cpp"):
        with open(os.path.join(ROOT, "ray_tracying_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


class Ranks:
    """torch.distributed plumbing of one bench process (rank)."""

    def __init__(self, gpus: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world == 1 and gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
        torch.cuda.set_device(self.local_rank)
        self.host_group = None
        if self.world > 1:
            # NCCL prints its version banner on stdout when the communicator is created: stdout carries the JSON line
            # and nothing else, so fd 1 points at stderr until the first collective is through
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
            # a HOST barrier (gloo): ranks that wait for rank 0's rt_render_multi must not spin on their GPUs
            self.host_group = dist.new_group(backend="gloo")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def max_(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_(self, values):
        t = self.torch.tensor(values, dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return [int(x) for x in t.tolist()]


def measure(rk: Ranks, name: str, steps: int, warmup: int, full: bool, tile=(32, 32), tile_block: int = 0) -> dict:
    """One workload on this job's ranks. full = also the roofline leg, the traversal ceiling and the clocks."""
    import ray_tracying_b200 as rt
    from ray_tracying_b200 import dist as rdist
    torch = rk.torch
    wl = WORKLOADS[name]
    R = wl["render"]
    world, rank = rk.world, rk.rank

    if rank == 0:
        scene_path_for(name)
    rk.host_barrier()
    path = scene_path_for(name)
    t0 = time.perf_counter()
    scene = rt.Scene.from_json(path, os.path.join(ROOT, "tests", "golden"))
    load_s = time.perf_counter() - t0
    width, height = scene.resolution
    common = dict(tile=tile, tile_block=tile_block, seed=1, **R)
    h2d_bytes = scene.upload()

    rgb = torch.zeros((height, width, 3), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()

    # one untimed pass with counters: rays per step and the traversal work (for the roofline)
    st = scene.render_device(rt.make_params(rank=rank, world=world, collect_stats=True, **common), rgb.data_ptr(), 0, 0, stream.cuda_stream)
    rays, n_primary, n_shadow, n_secondary, box_tests, prim_tests = rk.sum_(
        [st.rays, st.primary_rays, st.shadow_rays, st.secondary_rays, st.node_visits, st.prim_tests])
    my_box_tests = st.node_visits

    params = rt.make_params(rank=rank, world=world, **common)

    def one_step():
        scene.render_device(params, rgb.data_ptr(), 0, 0, stream.cuda_stream, sync_stats=False)
        return rdist.gather_frame(rgb, width, height, tile, rank, world, block=tile_block) if world > 1 else rgb

    clocks = ClockSampler(rk.local_rank)
    step_ms, kernel_ms = [], []
    launches0 = scene.launch_count()
    for i in range(warmup + steps):
        flush.zero_()  # evict the scene from L2 between iterations
        if i == warmup:
            launches0 = scene.launch_count()
            if rank == 0 and full:
                clocks.start()
        rk.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step()
        e1.record()
        rk.barrier()
        t = rk.max_([e0.elapsed_time(e1), scene.last_timing()[0]])
        if i >= warmup:
            step_ms.append(t[0])
            kernel_ms.append(t[1])
    launches = rk.sum_([scene.launch_count() - launches0])[0]  # kernels of librt_b200.so launched inside the timed region, all ranks
    clock_info = clocks.stop() if (rank == 0 and full) else {}
    ms = float(np.mean(step_ms))
    out = {
        "name": name, "desc": wl["desc"], "render": R, "resolution": [width, height], "shapes": scene.counts()["shapes"],
        "rays": rays, "ray_classes": {"primary": n_primary, "shadow": n_shadow, "secondary": n_secondary},
        "ms_per_step": ms, "value": rays / (ms * 1e-3) / 1e6, "kernel_ms_per_step": float(np.mean(kernel_ms)),
        "gpu_launches": launches, "load_s": load_s, "clocks": clock_info, "h2d_bytes": int(h2d_bytes),
        "box_tests": box_tests, "prim_tests": prim_tests,
    }

    # roofline leg: the same frames with the launches serialised on one stream and CUDA events around EVERY
    # trace / shadow / shade / light launch, so each kernel's duration is its own (in the headline steps
    # shadow/light of a level overlap trace/shade of the next level)
    if full:
        params_serial = rt.make_params(rank=rank, world=world, time_kernels=True, serial=True, **common)
        trav_ms, serial_ms, class_ms, n_trav = [], [], {}, 0
        for i in range(max(2, min(steps, 5))):
            flush.zero_()
            rk.barrier()
            scene.render_device(params_serial, rgb.data_ptr(), 0, 0, stream.cuda_stream, sync_stats=False)
            rk.barrier()
            kt = scene.last_kernel_times()
            for cls, (ms_k, n_timed, n_all) in kt.items():
                assert n_timed == n_all, "every launch of the frame is timed"
                class_ms.setdefault(cls, []).append(ms_k)
            # this rank's box-test rate is what the ceiling bounds; ranks differ little (interleaved tiles)
            trav_ms.append(kt["trace"][0] + kt["shadow"][0])
            serial_ms.append(scene.last_timing()[0])
            n_trav = kt["trace"][2] + kt["shadow"][2]
        trav = float(np.mean(trav_ms))
        peak_closest, _ = scene.traversal_peak(any_hit=False)
        peak_any, _ = scene.traversal_peak(any_hit=True)
        out["roofline_leg"] = {"trav_ms": trav, "serial_ms": float(np.mean(serial_ms)), "n_trav_launches": n_trav,
                               "class_ms": {k: float(np.mean(v)) for k, v in class_ms.items()},
                               "my_box_tests": my_box_tests, "peak_closest": peak_closest, "peak_any": peak_any}

    # end to end through the C ABI with HOST buffers, every step: scene evicted -> H2D copy of the scene from
    # page-locked host memory, render, D2H copy of the frame into page-locked host memory.
    #   1 GPU : rt_render (Scene.render_into);
    #   N GPUs: rt_render_multi driven by rank 0 alone (one process, N devices); the other ranks wait on the host.
    host_frame = torch.empty((height, width, 3), dtype=torch.uint8, pin_memory=True)
    p_e2e = rt.make_params(**common)
    e2e_s = []
    e2e_warm = max(1, min(warmup, 2))
    del flush
    torch.cuda.empty_cache()
    rk.host_barrier()
    if rank == 0:
        flush0 = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(world)]
        for i in range(e2e_warm + steps):
            for f in flush0:
                f.zero_()
            scene.evict()
            for d in range(world):
                torch.cuda.synchronize(d)
            w0 = time.perf_counter()
            if world == 1:
                scene.render_into(p_e2e, host_frame.data_ptr())
            else:
                scene.render_multi_into(p_e2e, world, host_frame.data_ptr())
            w1 = time.perf_counter()
            if i >= e2e_warm:
                e2e_s.append(w1 - w0)
        del flush0
    rk.host_barrier()
    if rank == 0:
        e2e_ms = float(np.mean(e2e_s)) * 1e3
        out["e2e"] = {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d_bytes) * world,
                      "d2h_bytes_per_step": int(width * height * 3), "ms_per_step": e2e_ms,
                      "path": "rt_render (C ABI, page-locked host buffers)" if world == 1 else
                              f"rt_render_multi (C ABI, one process driving {world} GPUs, page-locked host buffers, no NCCL)"}
    scene.close()
    torch.cuda.empty_cache()
    return out


def run_ours(args, wl):
    import ray_tracying_b200 as rt
    if rt.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    rk = Ranks(args.gpus)
    world, rank = rk.world, rk.rank
    tile = (32, 32)
    head = measure(rk, args.workload, args.steps, args.warmup, full=True, tile=tile, tile_block=args.tile_block)
    second = None
    if args.secondary and args.secondary != args.workload:
        second = measure(rk, args.secondary, args.steps, args.warmup, full=True, tile=tile, tile_block=args.tile_block)

    if rank == 0:
        tag = kernel_source_tag()
        captured = {}
        tp = os.path.join(ROOT, "profiles", "traffic_r2.json")
        if os.path.exists(tp):
            with open(tp) as f:
                captured = json.load(f)

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

        def roofline(m):
            # Dominant kernels = the traversal loops (trace_kernel / shadow_kernel / their packet flavours): one
            # number for all of them, box tests per second, against the measured ceiling of the loop's node step.
            leg = m["roofline_leg"]
            achieved = leg["my_box_tests"] / (leg["trav_ms"] * 1e-3) / 1e9
            peak = max(leg["peak_closest"], leg["peak_any"]) / 1e9
            cap = captured.get(m["name"]) if captured.get("kernel_source_tag") == tag else None
            hbm = None
            if cap:  # what the dominant kernel asks of HBM, against the driver-measured copy peak: a few % -- not the bound
                dram_gbs = cap["dram_bytes_per_launch"] / (cap["duration_us"] * 1e-6) / 1e9
                hbm = {"kernel": cap["kernel"], "dram_gbs": dram_gbs, "peak_gbs": hbm_peak, "frac": dram_gbs / hbm_peak, "peak_source": hbm_src}
            return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "Gbox-tests/s", "frac": achieved / peak,
                    "traffic": cap["dram_bytes_per_launch"] if cap else None,
                    "traffic_source": (captured.get("source") if cap else "no ncu capture of these kernel sources under profiles/"),
                    "hbm": hbm,
                    "peak_source": "rt_traversal_peak measured in this run: the loop's node step (4 slab tests, sorting network, 3 pushes, "
                                   "descent = full work) with converged warps on 64 L1-resident synthetic nodes; closest-hit %.1f / any-hit "
                                   "%.1f Gbox-tests/s" % (leg["peak_closest"] / 1e9, leg["peak_any"] / 1e9),
                    "kernel": "traversal loops: trace_packet_kernel + trace_kernel + shadow_packet_kernel + shadow_kernel",
                    "launches_per_step": leg["n_trav_launches"], "avg_launch_ms": leg["trav_ms"] / max(1, leg["n_trav_launches"]),
                    "box_tests_per_launch": leg["my_box_tests"] // max(1, leg["n_trav_launches"]),
                    "share_of_step": leg["trav_ms"] / leg["serial_ms"], "serialised_step_ms": leg["serial_ms"],
                    "kernel_class_ms_per_step": leg["class_ms"],
                    "timing": "CUDA events around every launch, launches serialised on one stream (roofline leg), rank 0's shard",
                    "box_tests_per_ray": m["box_tests"] / max(m["rays"], 1), "prim_tests_per_ray": m["prim_tests"] / max(m["rays"], 1),
                    "note": "the scene is L1/L2 resident (DRAM traffic is a few % of the bytes the box tests read), so the bound is "
                            "instruction issue, not HBM: frac = lanes doing box tests x issue rate relative to a loop with no divergence, "
                            "no cache misses and no fetch / primitive phases; primitive tests are not counted in `achieved`"}

        def block(m):
            return {"workload": m["desc"], "name": m["name"], "value": m["value"], "unit": "Mrays/s", "ms_per_step": m["ms_per_step"],
                    "rays_per_step": m["rays"], "rays": m["ray_classes"], "kernel_ms_per_step": m["kernel_ms_per_step"],
                    "e2e": m["e2e"], "roofline": roofline(m), "gpu_launches": m["gpu_launches"], "clocks": m["clocks"],
                    "host": {"scene_load_and_bvh_build_s": m["load_s"]}, **{k: v for k, v in m["render"].items()}}

        R = head["render"]
        line = {
            "metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": head["desc"], "name": head["name"], **R, "parallelism": f"tiles{world}", "tile": list(tile),
                       "tile_block": args.tile_block, "l2": "flushed between steps (256 MiB memset)", "shapes": head["shapes"],
                       "resolution": head["resolution"]},
            "rays_per_step": head["rays"], "rays": head["ray_classes"],
            "kernel_ms_per_step": head["kernel_ms_per_step"], "kernel_mrays_per_s": head["rays"] / (head["kernel_ms_per_step"] * 1e-3) / 1e6,
            "gpu_launches": head["gpu_launches"],
            "e2e": head["e2e"],
            "roofline": roofline(head),
            "clocks": head["clocks"],
            "host": {"scene_load_and_bvh_build_s": head["load_s"]},
            "kernel_source_tag": tag,
        }
        if second is not None:
            line["secondary"] = block(second)
        if world == 1 and not args.no_cpu_baseline:
            try:
                res = reference_sample(head["name"], scene_path_for(head["name"]), R, repeats=2, target_seconds=3.0)
                cms = float(np.mean(res["seconds"][1:])) * 1e3
                line["cpu_baseline"] = {"value": res["rays"] / (cms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": res["cores"], "kind": res["kind"],
                                        "sample": sample_text(res)}
            except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        print(json.dumps(line))
    rk.host_barrier()
    if world > 1:
        rk.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="soup1m")
    ap.add_argument("--secondary", default="mixed100k", help="second workload reported in the same line ('' = none)")
    ap.add_argument("--tile-block", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
