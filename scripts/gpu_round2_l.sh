#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2l_ab.jsonl; : > $OUT
run() { env "$@" 2>>gpurun_out/r2l_err.log | tail -1 >> $OUT; }
for wl in soup1m mixed100k glossy250k; do
  steps=8; [ $wl != mixed100k ] && steps=3
  run python scripts/perf_probe.py $wl $steps order_default
  run RT_B200_CHILD_ORDER=area python scripts/perf_probe.py $wl $steps order_area
  run RT_B200_CHILD_ORDER=small python scripts/perf_probe.py $wl $steps order_small
done
cat $OUT
