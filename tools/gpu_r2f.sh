#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
timeout 600 python tools/sweep.py config4 1 "" "split_k=8" "split_k=16" "split_k=32" "direct_len=14000,split_k=16" "direct_len=0,split_k=16" > gpurun_out/sweep_config4.txt 2>&1; cat gpurun_out/sweep_config4.txt
for q in 5000 10000 20000 35213; do timeout 120 python tools/c4_lone.py $q split_k=16; done 2>&1 | tee gpurun_out/c4_lone_all.txt
timeout 900 python tools/debug_r2c.py exact_full > gpurun_out/debug_exact_full.txt 2>&1; grep -c ok gpurun_out/debug_exact_full.txt; grep MISMATCH gpurun_out/debug_exact_full.txt | head -3
timeout 600 python tools/sweep.py config2 0.25,0.125 "" > gpurun_out/sweep_small.txt 2>&1; cat gpurun_out/sweep_small.txt
