#!/bin/bash
# round 2, call T (1 GPU): compaction reverted, wide-select fast path = histogram of chunk minima + 128-bit loads (opt-in): tests with
# the fast path forced on, A/B on the C3 bench, launch lists
mkdir -p gpurun_out
KB2_SELECT_FAST=1 timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/t_pytest_gpu.log 2>&1; echo "pytest(fast select) exit $?" | tee -a gpurun_out/t_pytest_gpu.log; tail -3 gpurun_out/t_pytest_gpu.log
run() { echo "--- $1"; env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/t_bench.err | tee -a gpurun_out/t_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'surv', j['roofline'].get('survivors_re_evaluated'), 'redone', j['roofline'].get('queries_redone'), 'clk', j['clocks'])"; }
run "KB2_SELECT_FAST=0"
run "KB2_SELECT_FAST=1"
run "KB2_SELECT_FAST=0"
run "KB2_SELECT_FAST=1"
KB2_SELECT_FAST=1 KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches_fast1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t_ncu_bench.log 2>&1; echo "ncu list exit $?"
KB2_SELECT_FAST=0 KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches_fast0.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t_ncu_bench0.log 2>&1; echo "ncu list exit $?"
for w in ivf_flat_1m ivf_pq_1m; do
  KB2_SELECT_FAST=1 timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['metric'], round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'))"
done
