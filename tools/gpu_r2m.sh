#!/bin/bash
# 1-GPU box: parity of the blocked one-lane order + swb_align_batch, then A/B of the pass-group / column-block sizes on the
# benchmark workload, the affine workload and the per-rank workloads of the 8-GPU layouts
mkdir -p gpurun_out
P=$PWD/ece1782-smith-waterman-cuda_b200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_tests.log
{
for L in lib_pg1 lib lib_pg4c32 lib_pg8c8 lib_pg8c16 lib_pg16c8 lib_pg2c32; do
  echo "== $L"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=3 timeout 400 python tools/sweep.py config2 1.0 "" 2>&1 | tail -1
done
echo "== short queries"
for L in lib_pg1 lib lib_pg8c8; do echo "== $L"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=2 timeout 400 python tools/sweep.py short 1.0 "" 2>&1 | tail -1; done
echo "== per-rank workloads of the 8-GPU layouts (P x R), old order vs new"
for L in lib_pg1 lib; do echo "== $L"; SWB_LIB=$P/$L/libswb.so SWEEP_REPS=3 timeout 600 python tools/sweep.py config2 1.0 "nshards=2,shard=0,qgroups=4,qgroup=0" "nshards=2,shard=1,qgroups=4,qgroup=3" "nshards=8,shard=0" "qgroups=8,qgroup=0" "nshards=2,shard=0,qgroups=4,qgroup=0,chunk_rows=2048" "nshards=2,shard=0,qgroups=4,qgroup=0,chunk_rows=3072" "nshards=2,shard=0,qgroups=4,qgroup=0,group_len=384" "nshards=2,shard=0,qgroups=4,qgroup=0,group_len=1536" 2>&1 | tail -8; done
} > gpurun_out/r2m_sweep.txt 2>&1
cat gpurun_out/r2m_sweep.txt | cut -c1-230
for L in lib_pg1 lib; do SWB_LIB=$P/$L/libswb.so timeout 300 python bench.py --steps 3 --warmup 2 --affine 10,2 --no-ref-cuda --no-cpu --e2e-steps 0 2>/dev/null | cut -c1-200; done > gpurun_out/r2m_affine.txt; cat gpurun_out/r2m_affine.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r2m_bench_1gpu.json; tail -3 gpurun_out/r2m_bench_1gpu.err
