#!/bin/bash
# round 2, call U (1 GPU): survivor rows trimmed to their k' best inside exact_eval (opt-in KB2_EVAL_TRIM=1): tests with it on, A/B
mkdir -p gpurun_out
KB2_EVAL_TRIM=1 timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/u_pytest_gpu.log 2>&1; echo "pytest(trim) exit $?" | tee -a gpurun_out/u_pytest_gpu.log; tail -3 gpurun_out/u_pytest_gpu.log
run() { echo "--- $1"; env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/u_bench.err | tee -a gpurun_out/u_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'surv', j['roofline'].get('survivors_re_evaluated'), 'redone', j['roofline'].get('queries_redone'))"; }
run "KB2_EVAL_TRIM=0"
run "KB2_EVAL_TRIM=1"
run "KB2_EVAL_TRIM=0"
run "KB2_EVAL_TRIM=1"
run "KB2_EVAL_TRIM=1 KB2_SELECT_FAST=0"
KB2_EVAL_TRIM=1 KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches_trim1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/u_ncu_bench.log 2>&1; echo "ncu list exit $?"
KB2_EVAL_TRIM=1 timeout 400 python bench.py --workload ivf_pq_1m --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['metric'], round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'), j['cpu_baseline'].get('parity_vs_gpu'))"
