#!/bin/bash
# 1-GPU box: the V16 cell with the additions beside the ALU pipe (SWB_V16_FORM 1: prmt, 2 vadd2, vimax3.relu, 1/2 vimax3)
# against the two-viaddmax cell (lib_form0), same GPU, same process order; then the GPU suite with the new cell
mkdir -p gpurun_out
PKG=ece1782-smith-waterman-cuda_b200
for rep in 1 2; do
SWB_LIB=$PWD/$PKG/lib_form0/libswb.so python tools/sweep.py config2 1.0 "" 2>&1 | sed 's/^/form0: /' | tee -a gpurun_out/r2x_sweep_v16_form.txt
python tools/sweep.py config2 1.0 "" 2>&1 | sed 's/^/form1: /' | tee -a gpurun_out/r2x_sweep_v16_form.txt
done
python tools/sweep.py short 1.0 "" 2>&1 | sed 's/^/form1 short: /' | tee -a gpurun_out/r2x_sweep_v16_form.txt
SWB_LIB=$PWD/$PKG/lib_form0/libswb.so python tools/sweep.py short 1.0 "" 2>&1 | sed 's/^/form0 short: /' | tee -a gpurun_out/r2x_sweep_v16_form.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_tests.log; tail -3 gpurun_out/r2x_tests.log
