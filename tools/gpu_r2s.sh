#!/bin/bash
# 1-GPU box: parity (affine traceback included), configs[3] and affine bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_tests.log; tail -3 gpurun_out/r2s_tests.log
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 --no-ref-cuda > gpurun_out/r2s_bench_config4.json 2> gpurun_out/r2s_bench_config4.err; cut -c1-200 gpurun_out/r2s_bench_config4.json
timeout 300 python bench.py --steps 3 --warmup 2 --affine 10,2 --no-ref-cuda > gpurun_out/r2s_bench_affine.json 2> gpurun_out/r2s_bench_affine.err; cut -c1-200 gpurun_out/r2s_bench_affine.json
