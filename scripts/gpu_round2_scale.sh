#!/bin/bash
# scaling lines for profiles/: bash scripts/gpu_round2_scale.sh "N1 N2 .." [steps] [extra]
#   per N: bench.py default (configs[2] + configs[1]) and configs[3] + configs[4]; extra=tb4: also tile blocks of 4 at the last N
NS=${1:-"1"}; STEPS=${2:-20}; EXTRA=${3:-}
mkdir -p gpurun_out
run_bench() {  # n, outfile, args...
  n=$1; out=$2; shift 2
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps $STEPS --warmup 3 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $STEPS --warmup 3 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  fi
  echo "$out rc=$?" >> gpurun_out/scale_rc.log
}
: > gpurun_out/scale_rc.log
for n in $NS; do
  run_bench $n scale_soup1m_mixed100k_n$n --no-cpu-baseline
  run_bench $n scale_glossy250k_dof4m_n$n --workload glossy250k --secondary dof4m --no-cpu-baseline
  last=$n
done
if [ "$EXTRA" = tb4 ]; then
  run_bench $last scale_tb4_soup1m_mixed100k_n$last --tile-block 4 --no-cpu-baseline
fi
cat gpurun_out/scale_rc.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_*_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    s = d.get('secondary', {})
    print(f"{f.split('/')[-1]:44s} N={d['n_gpus']} {d['config']['name']}: {d['value']:.0f} Mrays/s {d['ms_per_step']:.2f} ms e2e {d['e2e']['value']:.0f} frac {d['roofline']['frac']:.2f} | {s.get('name')}: {s.get('value', 0):.0f} Mrays/s {s.get('ms_per_step', 0):.3f} ms e2e {s.get('e2e', {}).get('value', 0):.0f} ({s.get('e2e', {}).get('ms_per_step', 0):.3f} ms) frac {s.get('roofline', {}).get('frac', 0):.2f}")
PY
