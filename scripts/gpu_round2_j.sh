#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2j_ab.jsonl; : > $OUT
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2j_err.log | tail -1 >> $OUT; }
for wl in mixed100k glossy250k dof4m soup1m; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps default
  for v in stream streamel; do
    run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py $wl $steps $v
  done
done
cat $OUT
