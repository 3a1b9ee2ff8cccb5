#!/bin/bash
# final single-GPU pass of round 2: full GPU test suite, smoke, the RT_CHECKS build (traps on traversal-stack overflow)
# on the tests and on the deepest tree, the default bench line, ncu captures of the final kernels
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/final_smoke.log
V=$PWD/ray_tracying_b200/variants
( time RT_B200_LIB=$V/librt_b200_checks.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x ) > gpurun_out/final_checks_pytest.log 2>&1
echo "checks pytest rc=$?" >> gpurun_out/final_checks_pytest.log
for wl in dof4m glossy250k; do RT_B200_LIB=$V/librt_b200_checks.so python scripts/perf_probe.py $wl 1 checks >> gpurun_out/final_checks_probe.jsonl 2>> gpurun_out/final_err.log; done
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?" >> gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
bash scripts/ncu_round2.sh > gpurun_out/final_ncu.log 2>&1
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/final_smoke.log | tail -2; tail -3 gpurun_out/final_checks_pytest.log; cat gpurun_out/final_checks_probe.jsonl | cut -c1-200; tail -2 gpurun_out/final_bench.err; head -c 600 gpurun_out/final_bench.json; echo; head -c 400 gpurun_out/final_bench_reference.json; echo; ls -la gpurun_out/prof_r2*.ncu-rep
