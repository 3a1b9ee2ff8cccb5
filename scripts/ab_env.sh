#!/bin/bash
# A/B of a run-time switch on the GPU box: bash scripts/ab_env.sh VAR "v1 v2" workload steps
# (RT_B200_OVERLAP, RT_B200_PACKET, RT_B200_PACKET_LEVELS, RT_B200_AREA_PACKETS, RT_B200_BATCH_SLOTS, RT_B200_SORT_EMIT)
VAR=$1; VALS=$2; WL=${3:-mixed100k}; STEPS=${4:-8}
for v in $VALS; do
  env $VAR=$v python scripts/perf_probe.py $WL $STEPS "$VAR=$v" 2>/dev/null | tail -1
done
