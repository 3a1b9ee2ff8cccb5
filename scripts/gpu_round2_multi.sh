#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): rt_render_multi tests, CLI -gpus, bench at 1..N ranks
N=${1:-2}; STEPS=${2:-5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or tile or shard" ) > gpurun_out/multi_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/multi_pytest.log
# the drop-in CLI over N GPUs against one GPU
python -c "
import sys; sys.path.insert(0, '.')
from ray_tracying_b200 import workloads
print(workloads.scene_path_for('mixed100k'))" > gpurun_out/multi_scene.txt 2>/dev/null
SCENE=$(tail -1 gpurun_out/multi_scene.txt)
./ray_tracying_b200/bin/Raytracer -input $SCENE -output /tmp/cli_1.ppm -bvh -s 1 -depth 5 -stats -textures tests/golden > gpurun_out/multi_cli_1.log 2>&1
./ray_tracying_b200/bin/Raytracer -input $SCENE -output /tmp/cli_n.ppm -bvh -s 1 -depth 5 -stats -gpus $N -textures tests/golden > gpurun_out/multi_cli_n.log 2>&1
cmp /tmp/cli_1.ppm /tmp/cli_n.ppm && echo "CLI: $N-GPU image identical to the 1-GPU image" >> gpurun_out/multi_cli_n.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/multi_bench_n$n.json 2> gpurun_out/multi_bench_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $STEPS --warmup 3 > gpurun_out/multi_bench_n$n.json 2> gpurun_out/multi_bench_n$n.err
  fi
  echo "bench n=$n rc=$?" >> gpurun_out/multi_pytest.log
done
tail -6 gpurun_out/multi_pytest.log; tail -3 gpurun_out/multi_cli_n.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/multi_bench_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    s = d.get('secondary', {})
    print(f"N={d['n_gpus']} {d['config']['name']}: {d['value']:.0f} Mrays/s {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.2f} ms) roofline {d['roofline']['frac']:.2f} | {s.get('name')}: {s.get('value', 0):.0f} Mrays/s {s.get('ms_per_step', 0):.3f} ms e2e {s.get('e2e', {}).get('value', 0):.0f} ({s.get('e2e', {}).get('ms_per_step', 0):.3f} ms)")
PY
