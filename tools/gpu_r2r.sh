#!/bin/bash
# 1-GPU box: LDS.64 profile loads (lib_ldw8) against the default build: parity of the variant, then A/B; configs[3] direct_len
mkdir -p gpurun_out
P=$PWD/ece1782-smith-waterman-cuda_b200
SWB_LIB=$P/lib_ldw8/libswb.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random or chunk or overflow or affine or shard" > gpurun_out/r2r_tests_ldw8.log 2>&1; tail -2 gpurun_out/r2r_tests_ldw8.log
{
for L in lib lib_ldw8 lib lib_ldw8; do echo "== $L"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=3 timeout 400 python tools/sweep.py config2 1.0 "" "group_len=384" 2>&1 | tail -2; done
for L in lib lib_ldw8; do echo "== $L short"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=2 timeout 400 python tools/sweep.py short 1.0 "" 2>&1 | tail -1; done
for L in lib lib_ldw8; do echo "== $L 1/8 shard"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=3 timeout 400 python tools/sweep.py config2 1.0 "nshards=8,shard=0" "nshards=2,shard=0,qgroups=4,qgroup=0" 2>&1 | tail -2; done
} > gpurun_out/r2r_sweep_ldw.txt 2>&1; cut -c1-200 gpurun_out/r2r_sweep_ldw.txt
SWEEP_REPS=6 timeout 600 python tools/sweep.py config4 1 "" "direct_len=16000" "direct_len=18000" "direct_len=21000" "" "direct_len=16000" "direct_len=18000" 2>&1 | cut -c1-200 | tee gpurun_out/r2r_sweep_config4.txt
