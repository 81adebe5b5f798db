#!/bin/bash
# GPU-box run: parity tests, then ncu on the bench command: launch list + one full capture of the score kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 30 -c 2 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | head -30
