#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2o_ab.jsonl; : > $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
run() { env "$@" 2>>gpurun_out/r2o_err.log | tail -1 >> $OUT; }
run python scripts/perf_probe.py soup1m 3 nosort
run python scripts/perf_probe.py mixed100k 8 nosort
tail -3 gpurun_out/r2o_pytest.log; cat $OUT | cut -c1-400
bash scripts/ncu_round2.sh > gpurun_out/r2o_ncu.log 2>&1
ls -la gpurun_out/prof_r2*.ncu-rep
