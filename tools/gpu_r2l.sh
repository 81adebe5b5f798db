#!/bin/bash
# 1-GPU box: full parity suite, the default bench line, the ncu launch list of the bench command and `--set full` captures of
# the kernels the bench runs now (each capture only after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_tests.log; tail -3 gpurun_out/r2l_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench_1gpu.json 2> gpurun_out/r2l_bench_1gpu.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r2l_bench_1gpu.json
Q="--steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled"
timeout 300 python bench.py $Q > gpurun_out/r2l_plain.log 2>&1 && {
ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv python bench.py $Q > gpurun_out/r2l_ncu_list.log 2>&1
$NCU -k regex:ILi32E3V16Li512ELi1ELb0E -c 1 -o gpurun_out/r2l_full_k32_512 python bench.py $Q > gpurun_out/r2l_ncu1.log 2>&1; tail -1 gpurun_out/r2l_ncu1.log | cut -c1-200
$NCU -k regex:ILi32E3V16Li256ELi2ELb0E -s 3 -c 1 -o gpurun_out/r2l_full_k32_256 python bench.py $Q > gpurun_out/r2l_ncu2.log 2>&1; tail -1 gpurun_out/r2l_ncu2.log | cut -c1-200
}
timeout 300 python bench.py $Q --scale 0.125 > gpurun_out/r2l_plain8.log 2>&1 &&
$NCU -k regex:ILi32E3V16Li256ELi2ELb0E -s 3 -c 2 -o gpurun_out/r2l_full_scale0125 python bench.py $Q --scale 0.125 > gpurun_out/r2l_ncu3.log 2>&1; tail -1 gpurun_out/r2l_ncu3.log | cut -c1-200
timeout 300 python bench.py $Q --affine 10,2 > gpurun_out/r2l_plain_aff.log 2>&1 &&
$NCU -k regex:ILi32E4V16A -s 3 -c 1 -o gpurun_out/r2l_full_affine python bench.py $Q --affine 10,2 > gpurun_out/r2l_ncu4.log 2>&1; tail -1 gpurun_out/r2l_ncu4.log | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 2 --affine 10,2 --no-ref-cuda > gpurun_out/r2l_bench_affine.json 2> gpurun_out/r2l_bench_affine.err; cut -c1-200 gpurun_out/r2l_bench_affine.json
timeout 120 python tools/c4_lone.py 35213 exact=1 > gpurun_out/r2l_c4_v32.log 2>&1 &&
$NCU -k regex:V32 -s 2 -c 1 -o gpurun_out/r2l_full_v32 python tools/c4_lone.py 35213 exact=1 > gpurun_out/r2l_ncu5.log 2>&1; tail -1 gpurun_out/r2l_ncu5.log | cut -c1-200
timeout 300 python bench.py --workload config4 --steps 5 --warmup 3 --no-ref-cuda > gpurun_out/r2l_bench_config4.json 2> gpurun_out/r2l_bench_config4.err; cut -c1-200 gpurun_out/r2l_bench_config4.json
ls -la gpurun_out/r2l_*
