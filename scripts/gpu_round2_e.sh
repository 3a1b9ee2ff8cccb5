#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2e_ab.jsonl; : > $OUT
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2e_err.log | tail -1 >> $OUT; }
for wl in mixed100k glossy250k soup1m; do
  steps=8; [ $wl != mixed100k ] && steps=3
  run python scripts/perf_probe.py $wl $steps default
  for v in fetch8 fetch16 spv2 spv4 stage21 stage85 anysortp; do
    [ $wl = soup1m ] && [ $v != anysortp ] && [ $v != stage21 ] && continue
    [ $wl != soup1m ] && [ $v = anysortp ] && continue
    run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py $wl $steps $v
  done
done
run RT_B200_PACKET_LEVELS=1 python scripts/perf_probe.py mixed100k 8 packets_l1
run RT_B200_BATCH_SLOTS=4194304 python scripts/perf_probe.py soup1m 3 batch4m
run RT_B200_BATCH_SLOTS=16777216 python scripts/perf_probe.py soup1m 3 batch16m
cat $OUT
