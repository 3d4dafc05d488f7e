#!/bin/bash
# round 2, call H (1 GPU): counting select in finalize, single-stage GEMM for short K, IVF_FLAT ticket draw, bound-kernel code cap,
# device-side HNSW construction (tests + build time), launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_hnsw_gpu.py tests/test_ivfpq_tc_gpu.py tests/test_ivf_gpu.py tests/test_flat_gpu.py tests/test_gemm_tc_gpu.py tests/test_golden_gpu.py tests/test_baseline_shapes_gpu.py -q -rf -x -s > gpurun_out/pytest_h.log 2>&1; echo "exit $?" >> gpurun_out/pytest_h.log; grep -E "hnsw recall|passed|failed|exit" gpurun_out/pytest_h.log | tail -8
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_h.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_h.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_GEMM_SHORTK=0"
run "KB2_TC_A_CODES=2000"
run "KB2_TC_A_CODES=1200"
for w in "" "KB2_TC_SCHED=static"; do
env $w timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('flat1m $w qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'stage', round(j['roofline']['scan_stage_ms'],3), 'recall', j['config']['recall_at_10'])"
done
for b in gpu host; do
KB2_HNSW_BUILD=$b timeout 900 python bench.py --workload hnsw_100k --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('hnsw100k build=$b qps', round(j['value']), 'recall', j['config'].get('recall_at_10'), 'build_s', j['config'].get('build_s'))"
done
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m_h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_h.log 2>&1
