#!/bin/bash
# 2-GPU box: the sharded load after the gather/plan overlap (db_load_ms of a half-database part), both layouts at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/r2q_box.log; nproc >> gpurun_out/r2q_box.log
timeout 600 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 3 --warmup 3 --db-parts 2 > gpurun_out/r2q_bench_2gpu_p2.json 2> gpurun_out/r2q_bench_2gpu_p2.err; echo "P=2 exit $?"
timeout 600 $TR --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err; echo "default exit $?"
python - <<'PY'
import json
for f in ('r2q_bench_2gpu_p2','r2q_bench_2gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), d['config']['layout']['db_parts'], 'e2e', round(d['e2e']['value']), 'load ms', round(d['e2e']['db_load_ms'],1), 'cold', round(d['e2e']['cold']['value']), d['sample_parity_ok'], d['topk_merge_ok'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/r2q_bench_2gpu_p2.err
