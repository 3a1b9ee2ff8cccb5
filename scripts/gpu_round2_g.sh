#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2g_ab.jsonl; : > $OUT
run() { env "$@" 2>>gpurun_out/r2g_err.log | tail -1 >> $OUT; }
for wl in mixed100k glossy250k dof4m; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps default
  run RT_B200_SORT_EMIT=1 python scripts/perf_probe.py $wl $steps sort_emit
done
cat $OUT
bash scripts/gpu_round2_scale.sh 1 20
P="python scripts/perf_probe.py"
$P soup1m 1 > gpurun_out/ncu_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'shadow_packet_kernel|trace_packet_kernel' -s 288 -c 2 -o gpurun_out/prof_r2_soup1m $P soup1m 1 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/prof_r2_soup1m.ncu-rep
