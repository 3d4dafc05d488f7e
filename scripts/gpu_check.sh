#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "exit $?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "
import json; j=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][0]); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']))"
