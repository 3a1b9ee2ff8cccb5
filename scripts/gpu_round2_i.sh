#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r2i_ab.jsonl; : > $OUT
V=$PWD/ray_tracying_b200/variants
run() { env "$@" 2>>gpurun_out/r2i_err.log | tail -1 >> $OUT; }
for wl in mixed100k glossy250k dof4m soup1m; do
  steps=8; [ $wl != mixed100k ] && steps=3; [ $wl = dof4m ] && steps=1
  run python scripts/perf_probe.py $wl $steps default
  for v in shade4 shade5 shade6; do
    run RT_B200_LIB=$V/librt_b200_$v.so python scripts/perf_probe.py $wl $steps $v
  done
done
cat $OUT
# memory checker on the small golden scenes (both traversal flavours, literal mode, linear scan, packets, windows)
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
   -k "(golden and (mixed_400 or numerics_edge or few_3 or textured)) or exhaustive or window or corners or random_streams" ) > gpurun_out/r2i_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2i_memcheck.log
tail -8 gpurun_out/r2i_memcheck.log
