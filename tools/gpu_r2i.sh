#!/bin/bash
mkdir -p gpurun_out
for L in lib_h0 lib_h16 lib_h32 lib_h64; do echo "== $L"; SWB_LIB=$PWD/ece1782-smith-waterman-cuda_b200/$L/libswb.so timeout 300 python tools/sweep.py config4 1 "split_k=16" "" 2>&1; done | tee gpurun_out/sweep_config4_hyst2.txt
