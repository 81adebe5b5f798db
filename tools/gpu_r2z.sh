#!/bin/bash
# 1-GPU box: L2-prefetch forms of the one-lane loop with the new V16 cell, same GPU (the variants lived behind a macro
# SWB_PF_MODE in commit da351ad: 0 top of the chunk, 1 none, 2 bulk prefetch per warp, 3 behind the columns = the code now)
mkdir -p gpurun_out
PKG=ece1782-smith-waterman-cuda_b200
for v in "" pf1 pf2 pf2s4 pf3 ""; do
  if [ -z "$v" ]; then L=$PWD/$PKG/lib/libswb.so; n=pf0; else L=$PWD/$PKG/lib_$v/libswb.so; n=$v; fi
  SWB_LIB=$L python tools/sweep.py config2 1.0 "" 2>&1 | sed "s/^/$n: /" | tee -a gpurun_out/r2z_sweep_pf_mode.txt
done
