#!/bin/bash
# final N=1 bench line of the committed sources + smoke
mkdir -p gpurun_out
timeout 240 python bench.py > gpurun_out/bench_r2_final_n1.json 2> gpurun_out/bench_r2_final_n1.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r2_final_n1.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
