#!/bin/bash
# round 2, call G (1 GPU): ticket scheduler of the filter kernel (A/B), refine-store fix, ncu --set full of the five top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ivfpq_tc_gpu.py tests/test_ivf_gpu.py tests/test_baseline_shapes_gpu.py -q -rf -x > gpurun_out/pytest_g.log 2>&1; echo "exit $?" >> gpurun_out/pytest_g.log; tail -4 gpurun_out/pytest_g.log
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_g.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_g.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_TC_SCHED=static"
run "KB2_TC_BALANCE=0"
run "KB2_TC_SCHED=static KB2_TC_BALANCE=0"
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:finalize_kernel|gemm_keys_tc_kernel|select_keys_kernel|bound_kernel|ivfpq_tc_filter_kernel|exact_eval_kernel|lut_build_kernel' \
  -o gpurun_out/top_kernels_g -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_g.log 2>&1
ls -la gpurun_out/top_kernels_g.ncu-rep
