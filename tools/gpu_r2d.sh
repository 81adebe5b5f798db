#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/debug_r2c.py exact_full > gpurun_out/debug_exact_full.txt 2>&1; cat gpurun_out/debug_exact_full.txt
for L in lib_r2a lib; do
  SWB_LIB=$PWD/ece1782-smith-waterman-cuda_b200/$L/libswb.so ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 1 -c 1 -o gpurun_out/grp_$L python tools/debug_r2c.py perf1 > gpurun_out/ncu_grp_$L.log 2>&1; tail -2 gpurun_out/ncu_grp_$L.log | cut -c1-200
done
