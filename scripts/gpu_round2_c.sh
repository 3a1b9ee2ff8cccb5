#!/bin/bash
# round 2, third GPU pass: the BVH4 over primitives (per-primitive gate) -- correctness, then A/B against the
# tree over reference leaves (variants/librt_b200_r2leaves.so) and occupancy variants
mkdir -p gpurun_out
OUT=gpurun_out/r2c_ab.jsonl; : > $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2c_err.log | tail -1 >> $OUT; }
for wl in mixed100k soup1m glossy250k; do
  steps=6; [ $wl != mixed100k ] && steps=3
  run python scripts/perf_probe.py $wl $steps new
  run RT_B200_LIB=$V/librt_b200_r2leaves.so python scripts/perf_probe.py $wl $steps r2leaves
done
for v in mb8 mb7 mb5; do
  run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py mixed100k 6 $v
  run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py soup1m 3 $v
done
tail -4 gpurun_out/r2c_pytest.log; cat $OUT
