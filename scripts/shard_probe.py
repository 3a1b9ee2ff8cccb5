"""Per-GPU cost of one shard of the default workload on ONE GPU (rank 0 of `world`): kernel time with and
without overlap, per-class serial times. Shows the fixed per-frame overhead that limits strong scaling."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import ray_tracying_b200 as rt

name = sys.argv[1] if len(sys.argv) > 1 else "mixed100k"
wl = bench.WORKLOADS[name]
scene = rt.Scene.from_json(bench.scene_path_for(name), os.path.join(bench.ROOT, "tests", "golden"))
w, h = scene.resolution
rgb = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for world in (1, 2, 4, 8):
    p = rt.make_params(rank=0, world=world, tile=(32, 32), seed=1, **wl["render"])
    ps = rt.make_params(rank=0, world=world, tile=(32, 32), seed=1, time_kernels=True, serial=True, **wl["render"])
    ks, ss = [], []
    for i in range(8):
        scene.render_device(p, rgb.data_ptr(), 0, 0, stream, sync_stats=False)
        torch.cuda.synchronize()
        ks.append(scene.last_timing()[0])
    for i in range(4):
        scene.render_device(ps, rgb.data_ptr(), 0, 0, stream, sync_stats=False)
        torch.cuda.synchronize()
        ss.append(scene.last_timing()[0])
    kt = scene.last_kernel_times()
    print(f"world={world}: kernel {min(ks[3:]):.3f} ms, serial {min(ss[1:]):.3f} ms, classes", {k: round(v[0], 3) for k, v in kt.items()})
