#!/bin/bash
# development run for the tensor-core PQ engine: its tests + the 10M bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ivfpq_tc_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc.log; tail -3 gpurun_out/pytest_tc.log
KB2_TC_VERBOSE=1 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep "kb2 tc" gpurun_out/bench_quick.err | tail -1
python -c "
import json; j=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][0]); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), j['clocks'])"
