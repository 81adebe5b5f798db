#!/bin/bash
# 1-GPU box: ncu --set full of the 256-thread bulk kernel with the HEAD cell: a mid-size query on the whole database,
# and the same kernel on the 1/8 part an 8-GPU run gives a rank
mkdir -p gpurun_out
SWEEP_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li256ELi2ELb0E -s 3 -c 1 -o gpurun_out/r2ze_full_k32_256 python tools/sweep.py config2 1.0 "" > gpurun_out/r2ze_ncu1.log 2>&1; tail -1 gpurun_out/r2ze_ncu1.log | cut -c1-200
SWEEP_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li256ELi2ELb0E -s 3 -c 1 -o gpurun_out/r2ze_full_eighth_k32_256 python tools/sweep.py config2 1.0 "nshards=8,shard=0" > gpurun_out/r2ze_ncu2.log 2>&1; tail -1 gpurun_out/r2ze_ncu2.log | cut -c1-200
ls -la gpurun_out/r2ze*
