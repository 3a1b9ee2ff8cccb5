#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2m_ab.jsonl; : > $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
run() { env "$@" 2>>gpurun_out/r2m_err.log | tail -1 >> $OUT; }
for wl in soup1m dof4m glossy250k mixed100k; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps light_ranks
  run RT_B200_NO_LIGHT_ORDER=1 python scripts/perf_probe.py $wl $steps slot_order
done
tail -4 gpurun_out/r2m_pytest.log; cat $OUT
