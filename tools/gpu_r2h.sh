#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep.py config4 1 "" "split_k=8" "split_k=16" "direct_len=14000,split_k=16" > gpurun_out/sweep_config4_even.txt 2>&1; cat gpurun_out/sweep_config4_even.txt
SWB_SPLIT_UNEVEN=1 timeout 600 python tools/sweep.py config4 1 "split_k=16" "direct_len=14000,split_k=16" 2>&1 | tee gpurun_out/sweep_config4_uneven.txt
for q in 5000 35213; do timeout 120 python tools/c4_lone.py $q split_k=16; done 2>&1 | tee gpurun_out/c4_lone_all.txt
