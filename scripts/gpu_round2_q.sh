#!/bin/bash
# configs[3] / configs[4] on one GPU with the final kernels (frame time + image checksum)
mkdir -p gpurun_out; OUT=gpurun_out/r2q_ab.jsonl; : > $OUT
timeout 70 python scripts/perf_probe.py glossy250k 3 final 2>>gpurun_out/r2q_err.log | tail -1 >> $OUT
timeout 100 python scripts/perf_probe.py dof4m 2 final 2>>gpurun_out/r2q_err.log | tail -1 >> $OUT
cut -c1-420 $OUT
