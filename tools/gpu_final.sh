#!/bin/bash
# GPU-box end-of-round check: smoke, parity suite, the default bench line, the reference arm, configs[3].
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err; echo "ref exit $?"; cut -c1-300 gpurun_out/final_bench_reference.json
timeout 300 python bench.py --workload config4 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/final_config4.json 2>> gpurun_out/final_bench.err; cut -c1-250 gpurun_out/final_config4.json
