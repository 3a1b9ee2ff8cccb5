#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2l2_ab.jsonl; : > $OUT
run() { env "$@" 2>>gpurun_out/r2l_err.log | tail -1 >> $OUT; }
for o in default low light_near light_far; do
  run RT_B200_CHILD_ORDER=$o python scripts/perf_probe.py soup1m 3 order_$o
done
for o in default low light_far; do
  run RT_B200_CHILD_ORDER=$o python scripts/perf_probe.py glossy250k 3 order_$o
done
cat $OUT
