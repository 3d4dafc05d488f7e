#!/bin/bash
# round 2, call W (1 GPU): long survivor rows trimmed too (staged), 256-thread phase A by default with the key buffer sized by the
# requested code count (4 CTAs/SM): all GPU tests, bench x3, launch list, C2 / 1M lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/w_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/w_pytest_gpu.log; tail -3 gpurun_out/w_pytest_gpu.log
run() { echo "--- $1"; env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/w_bench.err | tee -a gpurun_out/w_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'surv', j['roofline'].get('survivors_re_evaluated'), 'redone', j['roofline'].get('queries_redone'))"; }
run "KB2_NOOP=1"
run "KB2_EVAL_TRIM=0"
run "KB2_NOOP=1"
run "KB2_BOUND_NT=128"
KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/w_ncu_bench.log 2>&1; echo "ncu list exit $?"
for w in ivf_flat_1m ivf_pq_1m; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | tee -a gpurun_out/w_extra.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['metric'], round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'), (j.get('cpu_baseline') or {}).get('parity_vs_gpu'))"
done
