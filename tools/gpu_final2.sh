#!/bin/bash
# GPU-box end-of-round check (second half of round 2, new V16 cell): smoke, parity suite, the default bench line, the
# reference arm, configs[3], affine, and the ncu launch list of the bench command (the --set full capture of the same
# kernel is profiles/r2za_*)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final2_tests.log 2>&1; tail -2 gpurun_out/final2_tests.log
timeout 600 python bench.py --impl reference > gpurun_out/final2_bench_reference.json 2> gpurun_out/final2_bench.err; echo "ref exit $?"; cut -c1-300 gpurun_out/final2_bench_reference.json
timeout 600 python bench.py > gpurun_out/final2_bench.json 2>> gpurun_out/final2_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/final2_bench.json
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 --no-ref-cuda > gpurun_out/final2_config4.json 2>> gpurun_out/final2_bench.err; cut -c1-250 gpurun_out/final2_config4.json
timeout 300 python bench.py --affine 10,2 --steps 3 --warmup 3 --no-cpu --no-ref-cuda --e2e-steps 0 > gpurun_out/final2_affine.json 2>> gpurun_out/final2_bench.err; cut -c1-250 gpurun_out/final2_affine.json
Q="--steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/final2_launches.csv python bench.py $Q > gpurun_out/final2_ncu_list.log 2>&1; tail -1 gpurun_out/final2_ncu_list.log | cut -c1-200
