#!/bin/bash
# GPU-box check: parity tests, instruction-rate microbenchmarks, a small and the full bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.log 2>&1; nproc >> gpurun_out/box.log; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/box.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
timeout 120 python - > gpurun_out/microbench.log 2>&1 <<'PY'
import importlib
swb = importlib.import_module("ece1782-smith-waterman-cuda_b200")
for k, name in enumerate(swb.MICROBENCH_KINDS):
    print(name, round(swb.microbench(0, k), 1), "Glane-instr/s")
PY
timeout 600 python bench.py --steps 1 --warmup 1 --scale 0.1 --no-cpu --e2e-steps 1 > gpurun_out/bench_small.log 2>&1; echo "exit $?" >> gpurun_out/bench_small.log
timeout 1200 python bench.py --steps 2 --warmup 1 --per-query --cpu-seconds 8 > gpurun_out/bench_full.log 2>&1; echo "exit $?" >> gpurun_out/bench_full.log
tail -5 gpurun_out/tests.log; cat gpurun_out/microbench.log; tail -c 1500 gpurun_out/bench_small.log; tail -c 3000 gpurun_out/bench_full.log
