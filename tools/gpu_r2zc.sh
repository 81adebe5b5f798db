#!/bin/bash
# 1-GPU box: 20 warps per SM (640-thread blocks, 96 registers, a few spills) against 16 (512 threads, 128 registers)
# with the new V16 cell; K = 16 strips; whole database and the 1/8 part of an 8-GPU run
mkdir -p gpurun_out
PKG=ece1782-smith-waterman-cuda_b200
O=gpurun_out/r2zc_sweep_warps.txt
for v in "" nt640 ""  nt640; do
  if [ -z "$v" ]; then L=$PWD/$PKG/lib/libswb.so; n=nt512; else L=$PWD/$PKG/lib_$v/libswb.so; n=$v; fi
  SWB_LIB=$L python tools/sweep.py config2 1.0 "" "nshards=8,shard=0" 2>&1 | sed "s/^/$n: /" | tee -a $O
done
python tools/sweep.py config2 1.0 "k=16" 2>&1 | sed "s/^/nt512: /" | tee -a $O
