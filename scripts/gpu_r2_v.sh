#!/bin/bash
# round 2, call V (1 GPU): trim + fast select on by default, finalize tail pass as a row loop, 8 candidates in flight in the warp
# finalize, 256-thread phase A (opt-in KB2_BOUND_NT=256): tests (default and with the 256-thread phase A), A/B, launch lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/v_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/v_pytest_gpu.log; tail -3 gpurun_out/v_pytest_gpu.log
KB2_BOUND_NT=256 timeout 600 python -m pytest tests/test_ivfpq_tc_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q -x -rf > gpurun_out/v_pytest_nt256.log 2>&1; echo "pytest(nt256) exit $?" | tee -a gpurun_out/v_pytest_nt256.log; tail -2 gpurun_out/v_pytest_nt256.log
run() { echo "--- $1"; env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/v_bench.err | tee -a gpurun_out/v_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'surv', j['roofline'].get('survivors_re_evaluated'), 'redone', j['roofline'].get('queries_redone'))"; }
run "KB2_BOUND_NT=128"
run "KB2_BOUND_NT=256"
run "KB2_BOUND_NT=128"
run "KB2_BOUND_NT=256"
KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v_ncu_bench.log 2>&1; echo "ncu list exit $?"
KB2_BOUND_NT=256 KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches_nt256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v_ncu_bench2.log 2>&1; echo "ncu list exit $?"
