#!/bin/bash
# last pass of round 2: full GPU suite, smoke, default bench, then the ncu captures of exactly these sources
bash scripts/gpu_round2_verify.sh > gpurun_out/final2_verify.log 2>&1
bash scripts/ncu_round2.sh > gpurun_out/final2_ncu.log 2>&1
tail -12 gpurun_out/final2_verify.log | cut -c1-300; ls -la gpurun_out/prof_r2*.ncu-rep
