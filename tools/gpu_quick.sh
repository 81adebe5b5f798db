#!/bin/bash
# usage: gpu_quick.sh "<bench args 1>" "<bench args 2>" ...   -> one summary line per run
mkdir -p gpurun_out
i=0
for a in "$@"; do
  i=$((i+1))
  timeout 900 python bench.py --no-cpu $a > gpurun_out/quick_$i.log 2>&1; rc=$?
  python - "$a" gpurun_out/quick_$i.log $rc <<'PY'
import json,sys
args,path,rc=sys.argv[1],sys.argv[2],sys.argv[3]
ok=False
for l in open(path):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; ok=True
        print("[%s] value %.0f GCUPS  ms/step %.1f  e2e %.0f  frac %.3f" % (args, d["value"], d["ms_per_step"], d["e2e"]["value"] or 0, r["frac"]))
        if "per_query" in d: print("   ", {k:(round(v["gcups"]),round(v["ms"],2)) for k,v in d["per_query"].items()})
if not ok: print("[%s] FAILED rc=%s" % (args, rc)); print(open(path).read()[-1500:])
PY
done
