#!/bin/bash
# 8-GPU box: configs[4] (1,000 queries x UniProt-scale DB) and the configs[1]/[2] scaling points through bench.py, the
# engine group (one process, all GPUs), and the multi-device GPU tests on real devices
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box8.log 2>&1; nproc >> gpurun_out/box8.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -m gpu -x -q -k "engine_group or dropin_on_several" > gpurun_out/tests_8gpu.log 2>&1; tail -2 gpurun_out/tests_8gpu.log
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_config2_8gpu.json 2> gpurun_out/bench_config2_8gpu.err; echo "config2@8 exit $?"; cut -c1-330 gpurun_out/bench_config2_8gpu.json
timeout 600 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_config2_4gpu.json 2> gpurun_out/bench_config2_4gpu.err; echo "config2@4 exit $?"; cut -c1-330 gpurun_out/bench_config2_4gpu.json
timeout 600 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_config2_2gpu.json 2> gpurun_out/bench_config2_2gpu.err; echo "config2@2 exit $?"; cut -c1-330 gpurun_out/bench_config2_2gpu.json
timeout 600 $TR --nproc-per-node 8 --master-port 29614 bench.py --gpus 8 --steps 3 --warmup 3 --db-parts 8 > gpurun_out/bench_config2_8gpu_p8.json 2> gpurun_out/bench_config2_8gpu_p8.err; echo "config2@8 P=8 exit $?"; cut -c1-330 gpurun_out/bench_config2_8gpu_p8.json
timeout 600 python bench.py --gpus 8 --group --steps 3 --warmup 2 > gpurun_out/bench_config2_group8.json 2> gpurun_out/bench_config2_group8.err; echo "group@8 exit $?"; cut -c1-330 gpurun_out/bench_config2_group8.json
timeout 1500 $TR --nproc-per-node 8 --master-port 29615 bench.py --gpus 8 --workload config5 --steps 1 --warmup 1 > gpurun_out/bench_config5_8gpu.json 2> gpurun_out/bench_config5_8gpu.err; echo "config5@8 exit $?"; cut -c1-400 gpurun_out/bench_config5_8gpu.json; tail -c 400 gpurun_out/bench_config5_8gpu.err
