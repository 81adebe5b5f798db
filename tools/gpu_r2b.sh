#!/bin/bash
# round 2, second GPU call: restructured lane-group wavefront (parity suite again), long-sequence options, small-shard
# and P x R layout sweeps, ncu launch lists and full captures of the kernels the bench runs now
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 600 python tools/sweep.py config4 1 "" "split_k=8" "split_k=16" "split_k=32" "direct_len=14000,split_k=16" "direct_len=14000,split_k=32" "direct_len=0,split_k=16" "direct_len=0,exact=1" > gpurun_out/sweep_config4.txt 2>&1; cat gpurun_out/sweep_config4.txt
timeout 900 python tools/sweep.py config2 1.0 "" > gpurun_out/sweep_small.txt 2>&1
timeout 900 python tools/sweep.py config2 0.5 "" "split=1" "split=1,chunk_rows=2048" "qgroups=4,qgroup=0" "qgroups=4,qgroup=3" "qgroups=2,qgroup=1" "qgroups=4,qgroup=0,split=1" >> gpurun_out/sweep_small.txt 2>&1
timeout 900 python tools/sweep.py config2 0.25,0.125 "" "split=0" "group_len=384" >> gpurun_out/sweep_small.txt 2>&1
timeout 600 python tools/sweep.py config2 1.0 "qgroups=2,qgroup=0" "qgroups=4,qgroup=1" "qgroups=8,qgroup=0" "qgroups=8,qgroup=5" >> gpurun_out/sweep_small.txt 2>&1
cat gpurun_out/sweep_small.txt
# ncu: launch list of the bench command, then full captures (each after the plain command exited 0)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active --clock-control none -c 500 --csv --log-file gpurun_out/launches_config2.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -c 3 -o gpurun_out/r2_full_config2_long $CMD > gpurun_out/ncu_full1.log 2>&1; tail -2 gpurun_out/ncu_full1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 36 -c 3 -o gpurun_out/r2_full_config2_mid $CMD > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log | cut -c1-200
python tools/c4_lone.py 35213 > gpurun_out/c4_lone.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active --clock-control none -c 100 --csv --log-file gpurun_out/launches_c4_lone35213.csv python tools/c4_lone.py 35213 > gpurun_out/ncu_list_c4.log 2>&1
cat gpurun_out/c4_lone.log
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -c 3 -o gpurun_out/r2_full_c4_lone python tools/c4_lone.py 35213 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log | cut -c1-200
ls -la gpurun_out | head -40
