#!/bin/bash
# 1-GPU box: parity, then A/B of the block-coherent first wave (option static_wave) on the benchmark and on the per-rank
# workloads of the 8-GPU layouts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_tests.log; tail -3 gpurun_out/r2o_tests.log
SWEEP_REPS=3 timeout 900 python tools/sweep.py config2 1.0 "" "static_wave=0" "nshards=2,shard=0,qgroups=4,qgroup=0" "nshards=2,shard=0,qgroups=4,qgroup=0,static_wave=0" "nshards=2,shard=1,qgroups=4,qgroup=3" "nshards=2,shard=1,qgroups=4,qgroup=3,static_wave=0" "nshards=8,shard=0" "nshards=8,shard=0,static_wave=0" "nshards=4,shard=0,qgroups=2,qgroup=0" "nshards=4,shard=0,qgroups=2,qgroup=0,static_wave=0" "qgroups=4,qgroup=1" "qgroups=4,qgroup=1,static_wave=0" "" "static_wave=0" > gpurun_out/r2o_sweep.txt 2>&1; cut -c1-220 gpurun_out/r2o_sweep.txt
SWEEP_REPS=2 timeout 600 python tools/sweep.py short 1.0 "" "static_wave=0" 2>&1 | cut -c1-200 | tee -a gpurun_out/r2o_sweep.txt
SWEEP_REPS=3 timeout 300 python tools/sweep.py config4 1 "" 2>&1 | cut -c1-200 | tee -a gpurun_out/r2o_sweep.txt
