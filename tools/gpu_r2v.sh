#!/bin/bash
# 1-GPU box: parity with the word-load pack kernel, its steady-state time, the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2v_tests.log; tail -3 gpurun_out/r2v_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/r2v_pack_time.txt
import importlib, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
swb = importlib.import_module(bench.PKG)
codes, offsets = bench.synth_db(scale=1.0)
e = swb.Engine(0)
e.db_load(codes, offsets)
for _ in range(3):
    us, nb = e.pack_time(10)
    print("pack kernel: %.1f us per launch, %d bytes -> %.0f GB/s" % (us, nb, nb / us * 1e-3))
e.db_load(codes, offsets, 0, 8)
us, nb = e.pack_time(10); print("1/8 shard: %.1f us, %.0f GB/s" % (us, nb / us * 1e-3))
e.close()
PY
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_1gpu.json').read().strip().split('\n')[-1])
print(d['value'], d['e2e']['value'], d['e2e']['db_load_ms'], d['roofline']['frac'], d['roofline']['hbm_pack'], d['sample_parity_ok'], d['topk_merge_ok'])
PY
