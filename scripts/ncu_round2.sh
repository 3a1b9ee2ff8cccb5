#!/bin/bash
# ncu evidence for the round-2 kernels (one GPU; each ncu command directly after the same command ran clean):
#   1. launch list of one mixed100k frame + one soup1m frame (shares per kernel)
#   2. --set full of the traversal kernels + shade_kernel on mixed100k (level 0 packets, level 1 per-ray)
#   3. --set full of the dominant kernels of the headline workload (packet kernels on soup1m: level 0 of a batch in the
#      middle of the frame -- the first batches are sky)
mkdir -p gpurun_out
P="python scripts/perf_probe.py"
$P mixed100k 1 > gpurun_out/ncu_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r2_mixed100k.csv $P mixed100k 1 > gpurun_out/ncu_l1.log 2>&1
$P mixed100k 1 > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|shadow_kernel|trace_packet_kernel|shadow_packet_kernel|shade_kernel' -s 18 -c 6 -o gpurun_out/prof_r2_mixed100k $P mixed100k 1 > gpurun_out/ncu_f1.log 2>&1
$P soup1m 1 > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'shadow_packet_kernel|trace_packet_kernel' -s 288 -c 2 -o gpurun_out/prof_r2_soup1m $P soup1m 1 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 4. the shared-memory staging A/B (RT_STAGE_TOP=21 build): the per-ray kernels of level 1, to set L1 hit rate and
#    long_scoreboard against the default build's (capture 2)
export RT_B200_LIB=$PWD/ray_tracying_b200/variants/librt_b200_stage21.so
$P mixed100k 1 > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:'trace_kernel|shadow_kernel' -s 10 -c 2 -o gpurun_out/prof_r2_mixed100k_stage21 $P mixed100k 1 > gpurun_out/ncu_f4.log 2>&1
unset RT_B200_LIB
ls -la gpurun_out/*.ncu-rep
