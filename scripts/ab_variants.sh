#!/bin/bash
# A/B of kernel variants on the GPU box: for every ray_tracying_b200/variants/*.so (built here with
# `make -C ray_tracying_b200/csrc variant NAME=.. NVCC_EXTRA=..`) run bench.py and collect the lines.
# usage (under gpurun): bash scripts/ab_variants.sh [workload] [steps]
WL=${1:-mixed100k}; STEPS=${2:-10}
mkdir -p gpurun_out
OUT=gpurun_out/ab_${WL}.jsonl; : > $OUT
for so in ray_tracying_b200/librt_b200.so ray_tracying_b200/variants/*.so; do
  [ -f "$so" ] || continue
  line=$(RT_B200_LIB=$PWD/$so python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab_err.log | tail -1)
  echo "{\"lib\": \"$(basename $so)\", \"line\": $line}" >> $OUT
  python - "$so" "$line" <<'PY'
import json,sys
d=json.loads(sys.argv[2]); print(f"{sys.argv[1]:60s} {d['value']:9.1f} Mrays/s  {d['ms_per_step']:8.3f} ms  e2e {d['e2e']['value']:8.1f}")
PY
done
