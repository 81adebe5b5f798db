#!/bin/bash
# 1-GPU box: ncu --set full of the bulk one-lane kernel with the new V16 cell and the prefetch after the columns (longest query of the 20)
mkdir -p gpurun_out
SWEEP_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li512ELi1ELb0E -c 1 -o gpurun_out/r2za_full_k32_512 python tools/sweep.py config2 1.0 "" > gpurun_out/r2za_ncu.log 2>&1; tail -2 gpurun_out/r2za_ncu.log | cut -c1-300
ls -la gpurun_out/r2za*
