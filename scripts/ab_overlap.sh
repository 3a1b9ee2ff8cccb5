for ov in 1 0; do RT_B200_OVERLAP=$ov python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap=$ov', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'serial', round(d['roofline']['serialised_step_ms'],3), d['kernel_class_ms_per_step'])"; done
