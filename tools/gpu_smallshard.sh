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
run s8_xl3072 --scale 0.125
run s8_xl1536 --scale 0.125 --xl-len 1536
run s8_xl768 --scale 0.125 --xl-len 768
run s4_xl3072 --scale 0.25
run s4_xl1536 --scale 0.25 --xl-len 1536
run s1_split1 --split 1
run s1_split1_xl6144 --split 1 --xl-len 6144
run s1_split0 --split 0
