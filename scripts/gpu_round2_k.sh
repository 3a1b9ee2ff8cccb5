#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2k_ab.jsonl; : > $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
run() { env "$@" 2>>gpurun_out/r2k_err.log | tail -1 >> $OUT; }
for wl in soup1m glossy250k dof4m mixed100k; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run RT_B200_OCC_CACHE=0 python scripts/perf_probe.py $wl $steps occ0
  run RT_B200_OCC_CACHE=1 python scripts/perf_probe.py $wl $steps occ1
done
run RT_B200_OCC_CACHE=2 python scripts/perf_probe.py mixed100k 8 occ2
tail -4 gpurun_out/r2k_pytest.log; cat $OUT
