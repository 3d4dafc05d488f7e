#!/bin/bash
# round 2, call L (1 GPU): minima-histogram wide select (A/B), finalize split at 256, C4 (HNSW 1M x 768, device-built graph)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ivf_gpu.py tests/test_flat_gpu.py tests/test_ivfpq_tc_gpu.py tests/test_baseline_shapes_gpu.py tests/test_golden_gpu.py -q -x -rf > gpurun_out/pytest_l.log 2>&1; echo "exit $?" >> gpurun_out/pytest_l.log; grep -E "passed|failed|exit|Error" gpurun_out/pytest_l.log | tail -5
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_l.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_l.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_SELECT=hist"
timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('flat1m qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'])"
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_l.log 2>&1
timeout 900 python bench.py --workload hnsw_1m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hnsw1m_l.json 2> gpurun_out/bench_hnsw1m_l.err; python -c "
import json; j=json.loads(open('gpurun_out/bench_hnsw1m_l.json').read()); print('hnsw1m qps', round(j['value']), 'recall', j['config'].get('recall_at_10'), 'build_s', j['config'].get('build_s'), 'roofline', j['roofline'].get('frac'))" || tail -5 gpurun_out/bench_hnsw1m_l.err
