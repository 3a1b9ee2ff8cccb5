#!/usr/bin/env python
"""One line of timings for one workload on one GPU (A/B runs on the GPU box; bench.py is the judged number).

    [RT_B200_LIB=variant.so] [RT_B200_TREE=..] [RT_B200_STACK_CAP=..] python scripts/perf_probe.py WORKLOAD [steps] [tag]

Prints: frame ms (device, L2 flushed), Mrays/s, serialised per-class ms, box / primitive tests per ray.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ray_tracying_b200 as rt  # noqa: E402
from ray_tracying_b200 import workloads  # noqa: E402


def main():
    name = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    tag = sys.argv[3] if len(sys.argv) > 3 else ""
    R = workloads.WORKLOADS[name]["render"]
    path = workloads.scene_path_for(name)
    t0 = time.perf_counter()
    scene = rt.Scene.from_json(path, os.path.join(ROOT, "tests", "golden"))
    load_s = time.perf_counter() - t0
    w, h = scene.resolution
    rgb = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    st = scene.render_device(rt.make_params(seed=1, collect_stats=True, **R), rgb.data_ptr(), 0, 0, stream)
    p = rt.make_params(seed=1, **R)
    ps = rt.make_params(seed=1, time_kernels=True, serial=True, **R)
    ms = []
    for i in range(2 + steps):
        flush.zero_()
        torch.cuda.synchronize()
        scene.render_device(p, rgb.data_ptr(), 0, 0, stream, sync_stats=False)
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(scene.last_timing()[0])
    cls = {}
    for i in range(max(2, steps // 2)):
        flush.zero_()
        torch.cuda.synchronize()
        scene.render_device(ps, rgb.data_ptr(), 0, 0, stream, sync_stats=False)
        torch.cuda.synchronize()
        for k, v in scene.last_kernel_times().items():
            cls.setdefault(k, []).append(v[0])
    m = float(np.mean(ms))
    out = {"workload": name, "tag": tag, "lib": os.path.basename(os.environ.get("RT_B200_LIB", "default")),
           "tree": os.environ.get("RT_B200_TREE", "sah"), "stack_cap": os.environ.get("RT_B200_STACK_CAP", "default"),
           "ms": round(m, 3), "ms_min": round(float(np.min(ms)), 3), "mrays": round(st.rays / m / 1e3, 1),
           "class_ms": {k: round(float(np.mean(v)), 3) for k, v in cls.items()},
           "box_per_ray": round(st.node_visits / st.rays, 2), "prim_per_ray": round(st.prim_tests / st.rays, 3),
           "load_s": round(load_s, 2), "checksum": int(rgb.to(torch.int64).sum().item())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
