#!/bin/bash
# A/B of kernel variants on the GPU box: for the default build and every ray_tracying_b200/variants/*.so (built here with
# `make -C ray_tracying_b200/csrc variant NAME=.. NVCC_EXTRA=..`) run scripts/perf_probe.py and collect its lines.
# usage (under gpurun): bash scripts/ab_variants.sh [workload] [steps]
WL=${1:-mixed100k}; STEPS=${2:-8}
mkdir -p gpurun_out
OUT=gpurun_out/ab_${WL}.jsonl; : > $OUT
python scripts/perf_probe.py $WL $STEPS default 2>>gpurun_out/ab_err.log | tail -1 | tee -a $OUT
for so in ray_tracying_b200/variants/*.so; do
  [ -f "$so" ] || continue
  v=$(basename $so .so); v=${v#librt_b200_}
  RT_B200_LIB=$PWD/$so python scripts/perf_probe.py $WL $STEPS $v 2>>gpurun_out/ab_err.log | tail -1 | tee -a $OUT
done
