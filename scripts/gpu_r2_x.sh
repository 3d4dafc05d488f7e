#!/bin/bash
# round 2, call X (1 GPU): host queries uploaded in 4 pieces on the side stream, coarse stage per piece: tests, e2e A/B (100 steps)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/x_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/x_pytest_gpu.log; tail -3 gpurun_out/x_pytest_gpu.log
run() { echo "--- $1"; env $1 timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2> gpurun_out/x_bench.err | tee -a gpurun_out/x_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), 'e2e_ms', round(1e7/j['e2e']['value'],4), 'equal', j['e2e']['results_equal_device_path'])"; }
run "KB2_H2D_OVERLAP=0"
run "KB2_H2D_OVERLAP=1"
run "KB2_H2D_OVERLAP=0"
run "KB2_H2D_OVERLAP=1"
