#!/bin/bash
# 1-GPU box: the 1/8 part of an 8-GPU run (pure database split, all 20 queries) with the new cell: group_len, streams
mkdir -p gpurun_out
O=gpurun_out/r2zd_sweep_eighth.txt
S="nshards=8,shard=0"
SWEEP_REPS=4 python tools/sweep.py config2 1.0 "$S" "$S,group_len=192" "$S,group_len=768" "$S,group_len=1536" "$S,streams=20" "$S,streams=24" "$S,streams=10" "$S,static_wave=1" "$S,static_wave=0" "$S" 2>&1 | tee -a $O
