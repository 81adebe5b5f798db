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
run s8_new --scale 0.125
run s8_new_s20 --scale 0.125 --streams 20
run s8_new_chunk2k --scale 0.125 --chunk-rows 2048
run s8_new_chunk3k --scale 0.125 --chunk-rows 3072
run s8_new_chunk2k_s20 --scale 0.125 --chunk-rows 2048 --streams 20
run s8_new_s12 --scale 0.125 --streams 12
run s4_new --scale 0.25
run s4_new_chunk2k --scale 0.25 --chunk-rows 2048
run s1_new
