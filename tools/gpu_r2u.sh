#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pack_time or golden or align" > gpurun_out/r2u_tests.log 2>&1; tail -2 gpurun_out/r2u_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench_1gpu.json').read().strip().split('\n')[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['hbm_pack'], d['roofline']['traffic'])
PY
tail -3 gpurun_out/r2u_bench_1gpu.err
