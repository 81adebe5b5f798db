#!/bin/bash
# 4-GPU box: the driver's N = 4 command with the HEAD of the round (new V16 cell)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 110 $TR --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --steps 3 --warmup 3 --no-ref-cuda > gpurun_out/r2zf_bench_config2_4gpu.json 2> gpurun_out/r2zf_bench_config2_4gpu.err; echo "exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2zf_bench_config2_4gpu.json').read().strip().split('\n')[-1])
    print(round(d['value']), d['config']['layout'], 'e2e', round(d['e2e']['value']), 'load ms', round(d['e2e']['db_load_ms'],1), d['sample_parity_ok'], d['topk_merge_ok'], d['roofline']['frac'])
except Exception as e: print('failed', e)
PY
tail -2 gpurun_out/r2zf_bench_config2_4gpu.err
