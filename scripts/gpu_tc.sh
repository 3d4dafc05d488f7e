#!/bin/bash
# development run for the tensor-core PQ engine: its tests, a quick probe, the 10M bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ivfpq_tc_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_tc.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc.log; tail -5 gpurun_out/pytest_tc.log
for e in tc; do echo "== engine $e"; KB2_TC_VERBOSE=1 KB2_PQ_ENGINE=$e timeout 200 python scripts/quick_bench.py 1e6 10000 1024 64 2>&1 | grep -E "refine_k|recall|equal|rror|kb2 tc" | tail -3; done
KB2_TC_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep "kb2 tc" gpurun_out/bench_quick.err | tail -1
python -c "
import json; j=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][0]); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']))"
