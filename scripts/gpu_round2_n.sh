#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2n_ab.jsonl; : > $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
run() { env "$@" 2>>gpurun_out/r2n_err.log | tail -1 >> $OUT; }
for wl in dof4m glossy250k mixed100k; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps axis_fwd_rev
  run RT_B200_CHILD_ORDER=default python scripts/perf_probe.py $wl $steps construction
done
run python scripts/perf_probe.py soup1m 3 light_far
tail -4 gpurun_out/r2n_pytest.log; cat $OUT
