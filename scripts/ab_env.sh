#!/bin/bash
# A/B of a run-time switch on the GPU box: bash scripts/ab_env.sh VAR "v1 v2" workload steps warmup
VAR=$1; VALS=$2; WL=${3:-mixed100k}; STEPS=${4:-10}; WARM=${5:-3}
for v in $VALS; do
  env $VAR=$v python bench.py --workload $WL --steps $STEPS --warmup $WARM --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$WL $VAR=$v', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), ' serial', round(d['roofline']['serialised_step_ms'],3), {k: round(x,3) for k,x in d['kernel_class_ms_per_step'].items()}, 'box/ray', round(d['roofline']['box_tests_per_ray'],1), 'prim/ray', round(d['roofline']['prim_tests_per_ray'],2))"
done
