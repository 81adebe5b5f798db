#!/bin/bash
# 1-GPU box: parity with the new pack kernel, its duration (ncu launch list), half-shard group_len A/B, configs[3] options
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_tests.log; tail -3 gpurun_out/r2p_tests.log
Q="--steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"swb_pack|swb_topk|swb_scatter|swb_profile" -c 12 --csv --log-file gpurun_out/r2p_small_kernels.csv python bench.py $Q > /dev/null 2>&1
grep -E "swb_pack" gpurun_out/r2p_small_kernels.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tr -d '"'
SWEEP_REPS=3 timeout 900 python tools/sweep.py config2 1.0 "nshards=2,shard=0,qgroups=4,qgroup=0" "nshards=2,shard=0,qgroups=4,qgroup=0,group_len=1536" "nshards=2,shard=1,qgroups=4,qgroup=3" "nshards=2,shard=1,qgroups=4,qgroup=3,group_len=1536" "nshards=2,shard=0,qgroups=4,qgroup=1" "nshards=2,shard=0,qgroups=4,qgroup=1,group_len=1536" "nshards=2,shard=0,qgroups=4,qgroup=2" "nshards=2,shard=0,qgroups=4,qgroup=2,group_len=1536" "nshards=2,shard=0" "nshards=2,shard=0,group_len=1536" > gpurun_out/r2p_sweep.txt 2>&1; cut -c1-200 gpurun_out/r2p_sweep.txt
SWEEP_REPS=5 timeout 600 python tools/sweep.py config4 1 "" "streams=4" "direct_len=12000" "direct_len=16000" "batch_order=1" 2>&1 | cut -c1-200 | tee gpurun_out/r2p_sweep_config4.txt
