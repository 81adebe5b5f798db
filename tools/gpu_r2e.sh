#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
timeout 900 python tools/debug_r2c.py exact_full > gpurun_out/debug_exact_full.txt 2>&1; cat gpurun_out/debug_exact_full.txt
for L in lib_nopf lib; do
  echo "== $L"; SWB_LIB=$PWD/ece1782-smith-waterman-cuda_b200/$L/libswb.so timeout 600 python tools/sweep.py config2 1.0,0.5 "" 2>&1 | tee -a gpurun_out/sweep_ab.txt
done
timeout 600 python tools/sweep.py config4 1 "" "split_k=8" "split_k=16" "split_k=32" "direct_len=14000,split_k=16" "direct_len=0,split_k=16" > gpurun_out/sweep_config4.txt 2>&1; cat gpurun_out/sweep_config4.txt
timeout 900 python tools/sweep.py config2 0.5 "split=1" "qgroups=4,qgroup=0" "qgroups=4,qgroup=3" "qgroups=2,qgroup=1" > gpurun_out/sweep_small.txt 2>&1
timeout 900 python tools/sweep.py config2 0.25,0.125 "" >> gpurun_out/sweep_small.txt 2>&1
timeout 600 python tools/sweep.py config2 1.0 "qgroups=2,qgroup=0" "qgroups=4,qgroup=1" "qgroups=8,qgroup=0" >> gpurun_out/sweep_small.txt 2>&1
cat gpurun_out/sweep_small.txt
