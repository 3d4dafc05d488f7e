#!/bin/bash
# round 2, call J (1 GPU): cooperative-tile filter kernel (tests, A/B, ncu full), 512-entry warp finalize, TMEM read microbenchmark
mkdir -p gpurun_out
./scripts/micro/ldtm_bw > gpurun_out/ldtm_bw.txt 2>&1; cat gpurun_out/ldtm_bw.txt
timeout 1200 python -m pytest tests/test_ivfpq_tc_gpu.py tests/test_ivf_gpu.py tests/test_baseline_shapes_gpu.py tests/test_golden_gpu.py -q -rf -x > gpurun_out/pytest_j.log 2>&1; echo "exit $?" >> gpurun_out/pytest_j.log; grep -E "passed|failed|exit|Error" gpurun_out/pytest_j.log | tail -8
run() { echo "--- $1"; env $1 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_j.err | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'recall', j['config']['recall_at_10'], 'e2e', round(j['e2e']['value']), 'flagged', j['roofline'].get('queries_redone'))"; grep "kb2 tc" gpurun_out/bench_j.err | tail -1; }
run "KB2_NOOP=1"
run "KB2_TC_COOP=0"
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:ivfpq_tc_filter_kernel|select_keys_hist_kernel|finalize_warp_kernel' \
  -o gpurun_out/top_kernels_j -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_j.log 2>&1
ls -la gpurun_out/top_kernels_j.ncu-rep
