#!/bin/bash
# GPU-box A/B: parity tests, then the full bench with each s16 policy, plus per-query numbers.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
for v in 1 0; do
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 2 --variant $v --per-query > gpurun_out/bench_v$v.log 2>&1; echo "exit $?" >> gpurun_out/bench_v$v.log
  python - <<PY
import json
for l in open("gpurun_out/bench_v$v.log"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("variant $v value %.0f GCUPS  ms/step %.1f  e2e %.0f  frac %.3f peak %.0f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["peak"], d["clocks"]))
        print({k:(round(v["gcups"]),v["k"],round(v["ms"],2)) for k,v in d["per_query"].items()})
        print(r["instr_rates_glane_per_s"])
PY
done
tail -3 gpurun_out/bench_v1.log | cut -c1-600
