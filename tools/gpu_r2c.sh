#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_r2c.py exact > gpurun_out/debug_exact.txt 2>&1; cat gpurun_out/debug_exact.txt
SWB_LIB=$PWD/ece1782-smith-waterman-cuda_b200/lib_r2a/libswb.so timeout 600 python tools/debug_r2c.py perf > gpurun_out/debug_perf_r2a.txt 2>&1; cat gpurun_out/debug_perf_r2a.txt
timeout 600 python tools/debug_r2c.py perf > gpurun_out/debug_perf_new.txt 2>&1; cat gpurun_out/debug_perf_new.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "engine_group or dropin_on_several" > gpurun_out/tests_group.log 2>&1; tail -3 gpurun_out/tests_group.log
