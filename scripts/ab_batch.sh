for cfg in "8388608 1" "8388608 0" "3145728 1" "3145728 0" "16777216 1"; do set -- $cfg; RT_B200_BATCH_SLOTS=$1 RT_B200_AREA_PACKETS=$2 python bench.py --workload soup1m --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('soup1m batch=$1 area_packets=$2', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],1), 'ms')"; done
