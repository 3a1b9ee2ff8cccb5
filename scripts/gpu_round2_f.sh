#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2f_ab.jsonl; : > $OUT
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2f_err.log | tail -1 >> $OUT; }
for wl in mixed100k glossy250k dof4m; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps default
  for so in $V/*.so; do
    v=$(basename $so .so); v=${v#librt_b200_}
    run RT_B200_LIB=$so python scripts/perf_probe.py $wl $steps $v
  done
done
run python scripts/perf_probe.py soup1m 3 default
run RT_B200_LIB=$V/librt_b200_wave8.so python scripts/perf_probe.py soup1m 3 wave8
cat $OUT
# the dominant kernels of the headline workload under ncu (level 0 of the first batch of the second frame)
P="python scripts/perf_probe.py"
$P soup1m 1 > gpurun_out/ncu_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'shadow_packet_kernel|trace_packet_kernel' -s 192 -c 2 -o gpurun_out/prof_r2_soup1m $P soup1m 1 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/prof_r2_soup1m.ncu-rep
