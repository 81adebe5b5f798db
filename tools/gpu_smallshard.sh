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
run s8_split_chunk2k --scale 0.125 --split 1 --chunk-rows 2048
run s8_split_xl4k --scale 0.125 --split 1 --xl-len 4096
run s8_split_xl2k --scale 0.125 --split 1 --xl-len 2048
run s4_split --scale 0.25 --split 1
run s2_lpt --scale 0.5
run s2_split --scale 0.5 --split 1
run s1_split --split 1
