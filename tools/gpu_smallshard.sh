#!/bin/bash
# GPU-box experiment: the per-GPU workload of the 8-GPU strong-scaling run (1/8 of configs[1]) on ONE GPU,
# with the engine options that could matter for small shards.
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 "$@" > gpurun_out/ss_$name.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/ss_$name.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print("$name: value %.0f GCUPS ms/step %.2f e2e %.0f launches %s tiles %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["engine"]["tiles_by_group"]))
PY
}
run gl384 --steps 2
run gl768 --steps 2 --group-len 768
run gl1536 --steps 2 --group-len 1536
run sq_gl384 --steps 2 --synth-queries 150
run sq_gl768 --steps 2 --synth-queries 150 --group-len 768
run sq_gl1536 --steps 2 --synth-queries 150 --group-len 1536
run s8_gl768 --scale 0.125 --group-len 768
