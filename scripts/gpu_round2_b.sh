#!/bin/bash
# round 2, second GPU pass: correctness of the SAH upper tree + short stacks, then A/B of tree / stack cap / staging
mkdir -p gpurun_out
OUT=gpurun_out/r2b_ab.jsonl; : > $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2b_err.log | tail -1 >> $OUT; }
for tree in median sah; do for cap in 64 24 16 12; do
  run RT_B200_TREE=$tree RT_B200_STACK_CAP=$cap python scripts/perf_probe.py mixed100k 8
done; done
for v in stage21 stage85 ovfcall; do
  run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py mixed100k 8 $v
done
for tree in median sah; do for cap in 64 16; do
  run RT_B200_TREE=$tree RT_B200_STACK_CAP=$cap python scripts/perf_probe.py soup1m 3
done; done
run RT_B200_LIB=$V/librt_b200_stage21.so python scripts/perf_probe.py soup1m 3 stage21
for tree in median sah; do
  run RT_B200_TREE=$tree python scripts/perf_probe.py glossy250k 2
done
tail -4 gpurun_out/r2b_pytest.log; cat $OUT
