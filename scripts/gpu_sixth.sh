#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -6 gpurun_out/pytest.log
for nt in 256 512; do
  KB2_SCAN_NT=$nt KB2_GEMM=tc timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_nt$nt.json 2> gpurun_out/bench_nt$nt.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_10m_nt$nt.json')); print('NT=$nt', 'qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']))"
done
timeout 900 python scripts/extra_bench.py all 2>&1 | grep -v WARNING | tee gpurun_out/extra_bench.jsonl
