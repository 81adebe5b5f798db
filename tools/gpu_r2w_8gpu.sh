#!/bin/bash
# 8-GPU box, lean: configs[1] at 8 GPUs with the final code (default layout and pure database split)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2w_bench_config2_8gpu.json 2> gpurun_out/r2w_bench_config2_8gpu.err; echo "config2@8 exit $?"
timeout 400 $TR --nproc-per-node 8 --master-port 29632 bench.py --gpus 8 --steps 3 --warmup 3 --db-parts 8 > gpurun_out/r2w_bench_config2_8gpu_p8.json 2> gpurun_out/r2w_bench_config2_8gpu_p8.err; echo "config2@8 P=8 exit $?"
python - <<'PY'
import json
for f in ('r2w_bench_config2_8gpu','r2w_bench_config2_8gpu_p8'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), d['config']['layout']['db_parts'], 'e2e', round(d['e2e']['value']), 'load ms', round(d['e2e']['db_load_ms'],1), 'cold', round(d['e2e']['cold']['value']), d['sample_parity_ok'], d['topk_merge_ok'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
PY
tail -2 gpurun_out/r2w_bench_config2_8gpu.err
