#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2h_ab.jsonl; : > $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2h_err.log | tail -1 >> $OUT; }
for wl in mixed100k soup1m glossy250k dof4m; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps hitfirst
  run RT_B200_LIB=$V/librt_b200_gatefirst.so python scripts/perf_probe.py $wl $steps gatefirst
done
tail -4 gpurun_out/r2h_pytest.log; cat $OUT
