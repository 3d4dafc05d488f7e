#!/bin/bash
# round 2, call O (1 GPU): level-wise wide select, multi-group phase A (m48): tests, C5 smoke at N=1, C3 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ivfpq_tc_gpu.py tests/test_ivf_gpu.py tests/test_flat_gpu.py tests/test_baseline_shapes_gpu.py tests/test_golden_gpu.py tests/test_fourcc_gpu.py -q -x -rf > gpurun_out/pytest_o.log 2>&1; echo "exit $?" >> gpurun_out/pytest_o.log; grep -E "passed|failed|exit|Error" gpurun_out/pytest_o.log | tail -5
KB2_TC_VERBOSE=1 timeout 900 python scripts/bench_c5.py --rows 10000000 --nlist 8192 --steps 5 --warmup 2 > gpurun_out/c5_smoke_o.json 2> gpurun_out/c5_smoke_o.err; echo "exit $?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/c5_smoke_o.json') if l.startswith('{')][0]); print('C5 10M N=1 qps', round(j['value']), 'ms', round(j['ms_per_step'],3), j['stage_breakdown_rank0_ms'], 'recall', j['config']['recall_at_10_pure_adc'])"; grep "kb2 tc" gpurun_out/c5_smoke_o.err | tail -1
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_o.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_o.err | tail -1; }
run "KB2_NOOP=1"
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_o.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_o.log 2>&1
