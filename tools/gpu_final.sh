#!/bin/bash
# GPU-box end-of-round check: smoke, parity suite, the default bench line, the reference arm, configs[3], affine, and the
# ncu evidence of the same bench command (launch list + one full capture of the dominant kernel)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err; echo "ref exit $?"; cut -c1-300 gpurun_out/final_bench_reference.json
timeout 600 python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/final_bench.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-ref-cuda > gpurun_out/final_bench_10steps.json 2>> gpurun_out/final_bench.err; cut -c1-200 gpurun_out/final_bench_10steps.json
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 --no-ref-cuda > gpurun_out/final_config4.json 2>> gpurun_out/final_bench.err; cut -c1-250 gpurun_out/final_config4.json
Q="--steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
timeout 300 python bench.py $Q > gpurun_out/final_plain.log 2>&1 && {
ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py $Q > gpurun_out/final_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li512ELi1ELb0E -c 1 -o gpurun_out/final_full_k32_512 python bench.py $Q > gpurun_out/final_ncu1.log 2>&1; tail -1 gpurun_out/final_ncu1.log | cut -c1-200
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li256ELi2ELb0E -s 3 -c 1 -o gpurun_out/final_full_k32_256 python bench.py $Q > gpurun_out/final_ncu2.log 2>&1; tail -1 gpurun_out/final_ncu2.log | cut -c1-200
}
