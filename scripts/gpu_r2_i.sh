#!/bin/bash
# round 2, call I (1 GPU): warp-per-query finalize (tests, A/B), TMEM read microbenchmark
mkdir -p gpurun_out
./scripts/micro/ldtm_bw > gpurun_out/ldtm_bw.txt 2>&1; cat gpurun_out/ldtm_bw.txt
timeout 1200 python -m pytest tests/test_ivf_gpu.py tests/test_flat_gpu.py tests/test_ivfpq_tc_gpu.py tests/test_hnsw_gpu.py tests/test_golden_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fourcc_gpu.py -q -rf -x > gpurun_out/pytest_i.log 2>&1; echo "exit $?" >> gpurun_out/pytest_i.log; grep -E "passed|failed|exit|Error" gpurun_out/pytest_i.log | tail -8
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_i.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_i.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_FINALIZE=cta"
timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('flat1m qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'])"
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_i.log 2>&1
