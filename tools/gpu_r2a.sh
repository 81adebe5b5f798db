#!/bin/bash
# round 2, first GPU call: parity suite, bench line, long-sequence options, group_len sweep on smaller shards
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.log 2>&1; nproc >> gpurun_out/box.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 900 python bench.py --steps 2 --warmup 2 --cpu-seconds 6 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
tail -c 600 gpurun_out/bench_1gpu.err; cut -c1-400 gpurun_out/bench_1gpu.json
timeout 600 python tools/sweep.py config4 1 "" "split_k=8" "split_k=16" "direct_len=0" "direct_len=0,exact=1" "direct_len=0,exact=1,split_k=8" "direct_len=6000" "direct_len=14000" > gpurun_out/sweep_config4.txt 2>&1; cat gpurun_out/sweep_config4.txt
timeout 900 python tools/sweep.py config2 1.0,0.5,0.25,0.125 "" "group_len=384" "group_len=768" "group_len=1536" > gpurun_out/sweep_group_len.txt 2>&1; cat gpurun_out/sweep_group_len.txt
timeout 300 python bench.py --workload config4 --steps 3 --warmup 2 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; cut -c1-300 gpurun_out/bench_config4.json; tail -c 300 gpurun_out/bench_config4.err
