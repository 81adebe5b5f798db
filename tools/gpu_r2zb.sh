#!/bin/bash
# 1-GPU box: knobs around the new V16 cell (LDS.64 profile loads, L2-prefetch distance) on configs[1]; the new cell
# against the old one (lib_form0) on configs[3] (V16R, pipelined passes) and with affine gaps (V16A)
mkdir -p gpurun_out
PKG=ece1782-smith-waterman-cuda_b200
O=gpurun_out/r2zb_sweeps.txt
for v in "" ldw8 pfc3 pfc12 ""; do
  if [ -z "$v" ]; then L=$PWD/$PKG/lib/libswb.so; n=default; else L=$PWD/$PKG/lib_$v/libswb.so; n=$v; fi
  SWB_LIB=$L python tools/sweep.py config2 1.0 "" 2>&1 | sed "s/^/$n: /" | tee -a $O
done
for v in form0 "" form0 ""; do
  if [ -z "$v" ]; then L=$PWD/$PKG/lib/libswb.so; n=form1; else L=$PWD/$PKG/lib_$v/libswb.so; n=$v; fi
  SWB_LIB=$L python tools/sweep.py config4 1.0 "" 2>&1 | sed "s/^/$n: /" | tee -a $O
done
for v in form0 ""; do
  if [ -z "$v" ]; then L=$PWD/$PKG/lib/libswb.so; n=form1; else L=$PWD/$PKG/lib_$v/libswb.so; n=$v; fi
  SWB_LIB=$L python bench.py --affine 10,2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$n affine 10,2:', round(d['value'],1), 'GCUPS', d['ms_per_step'], 'ms')" | tee -a $O
done
