#!/bin/bash
# 1-GPU box: parity, DRAM traffic of the 5,478-row launch per build (old order / 4 x 16 / 8 x 16), one full capture of the
# new default, bench lines (linear, affine A/B), sharded-load timing
mkdir -p gpurun_out
P=$PWD/ece1782-smith-waterman-cuda_b200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_tests.log; tail -3 gpurun_out/r2n_tests.log
Q="--steps 1 --warmup 1 --no-cpu --no-ref-cuda --e2e-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2n_plain.log 2>&1 && {
for L in lib_pg1 lib lib_pg8c16; do
SWB_LIB=$P/$L/libswb.so ncu --metrics gpu__time_duration.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base mangled -k regex:ILi32E3V16Li512ELi1ELb0E -c 2 --csv --log-file gpurun_out/r2n_traffic_$L.csv python bench.py $Q > /dev/null 2>&1
echo "== $L"; grep -E "dram__bytes|gpu__time|alu_cycles|hit_rate" gpurun_out/r2n_traffic_$L.csv | cut -d, -f5,13- | head -10
done
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ILi32E3V16Li512ELi1ELb0E -c 1 -o gpurun_out/r2n_full_k32_512 python bench.py $Q > gpurun_out/r2n_ncu1.log 2>&1; tail -1 gpurun_out/r2n_ncu1.log | cut -c1-200
}
for L in lib_pg1 lib lib_pg8c16; do echo "== $L"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=3 timeout 400 python tools/sweep.py config2 1.0 "" 2>&1 | tail -1; done > gpurun_out/r2n_sweep.txt 2>&1; cat gpurun_out/r2n_sweep.txt | cut -c1-200
for L in lib_pg1 lib; do SWB_LIB=$P/$L/libswb.so timeout 300 python bench.py --steps 3 --warmup 2 --affine 10,2 --no-ref-cuda --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-200; done > gpurun_out/r2n_affine.txt; cat gpurun_out/r2n_affine.txt
SWB_TIMING=1 SWEEP_REPS=1 timeout 300 python tools/sweep.py config2 1.0 "nshards=2,shard=0" "nshards=8,shard=3" "" 2>&1 | grep -E "db_load|GCUPS" | cut -c1-330 > gpurun_out/r2n_load_timing.txt; cat gpurun_out/r2n_load_timing.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r2n_bench_1gpu.json; tail -3 gpurun_out/r2n_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_1gpu.json').read().strip().split('\n')[-1]); print(d.get('align')); print(d['roofline']['frac'], d['e2e']['value'])
PY
