#!/bin/bash
# GPU-box run: parity tests, e2e timing breakdown, reference CUDA kernel beside ours, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
SWB_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/bench_timing.log 2>&1
grep "swb\]" gpurun_out/bench_timing.log | tail -4
# reference CUDA solver (unmodified SWSolver.cu, sm_100a) vs ours on the same files (quarter-size DB: the
# reference's fixed 400 MB residue buffer holds 2e8 shorts)
python tools/make_fasta.py /tmp/db_q.fasta 0.25 > gpurun_out/refcuda.log 2>&1
for q in P02232 P01008 P27895; do
  timeout 600 oracle/_ref/ref_cuda_scan tests/golden/queries/$q.fasta /tmp/db_q.fasta 2 > /tmp/ref_$q.txt 2>>gpurun_out/refcuda.log
  grep "#TIME" /tmp/ref_$q.txt >> gpurun_out/refcuda.log
  timeout 600 ece1782-smith-waterman-cuda_b200/bin/main --query tests/golden/queries/$q.fasta --db /tmp/db_q.fasta > /tmp/ours_$q.txt 2>>gpurun_out/refcuda.log
  tail -6 /tmp/ours_$q.txt >> gpurun_out/refcuda.log
  grep -v "^#" /tmp/ref_$q.txt | sort > /tmp/a.txt; grep -E "^-?[0-9]+:-?[0-9]+$" /tmp/ours_$q.txt | sort > /tmp/b.txt
  echo "$q identical id:score lines: $(comm -12 /tmp/a.txt /tmp/b.txt | wc -l) of $(wc -l < /tmp/a.txt) (ref) / $(wc -l < /tmp/b.txt) (ours)" >> gpurun_out/refcuda.log
done
cat gpurun_out/refcuda.log | tail -30
# ncu: launch list of the bench command (plain run first), then one full capture of the score kernel
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:swb_score_kernel -s 24 -c 2 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
