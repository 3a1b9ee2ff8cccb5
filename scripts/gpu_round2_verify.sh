#!/bin/bash
# re-verification after host-side changes: full GPU suite + smoke + a short default bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/verify_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/verify_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/verify_smoke.log
( time python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
echo "bench rc=$?" >> gpurun_out/verify_bench.err
RT_B200_DEBUG=1 python scripts/perf_probe.py dof4m 1 load > gpurun_out/verify_dof4m.json 2> gpurun_out/verify_dof4m.err
tail -3 gpurun_out/verify_pytest.log; tail -2 gpurun_out/verify_smoke.log; tail -2 gpurun_out/verify_bench.err; grep -E "load |flatten:" gpurun_out/verify_dof4m.err; cat gpurun_out/verify_dof4m.json | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/verify_bench.json').read().strip().splitlines()[-1])
r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['traffic'], r['hbm'])
s=d['secondary']; print(s['value'], s['ms_per_step'], s['e2e']['value'], s['roofline']['frac'], s['roofline']['hbm'])
PY
