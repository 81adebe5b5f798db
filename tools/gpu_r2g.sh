#!/bin/bash
mkdir -p gpurun_out
python tools/c4_lone.py 35213 split_k=16 > gpurun_out/c4_lone.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 6 -c 2 -o gpurun_out/r2g_full_c4_lone python tools/c4_lone.py 35213 split_k=16 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log | cut -c1-200
python tools/c4_lone.py 5000 split_k=16 > gpurun_out/c4_lone5k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 4 -c 1 -o gpurun_out/r2g_full_c4_lone5k python tools/c4_lone.py 5000 split_k=16 > gpurun_out/ncu_full4.log 2>&1; tail -2 gpurun_out/ncu_full4.log | cut -c1-200
cat gpurun_out/c4_lone.log gpurun_out/c4_lone5k.log
