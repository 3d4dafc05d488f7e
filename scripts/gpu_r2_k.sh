#!/bin/bash
# round 2, call K (1 GPU): adaptive warp finalize + packed final sort, thread-minima wide select: all GPU tests, bench, launch list
mkdir -p gpurun_out
for f in tests/test_*_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -x -rf > gpurun_out/pytest_k_$n.log 2>&1; echo "== $n: $(tail -1 gpurun_out/pytest_k_$n.log) exit $?"
done
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_k.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_k.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_FINALIZE=cta"
timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('flat1m qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'])"
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_k.log 2>&1
