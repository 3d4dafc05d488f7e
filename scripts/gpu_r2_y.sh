#!/bin/bash
# round 2, call Y (1 GPU): final validation of the round-2 state -- every GPU test, smoke(), default bench
# (with CPU baseline + id parity), ncu launch list, one --set full capture of the top kernels, the other workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/y_smi.txt
nproc > gpurun_out/y_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/y_host.txt 2>/dev/null
timeout 1100 python -m pytest tests -m gpu -q -x -rf --durations=15 > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/y_pytest_gpu.log; tail -3 gpurun_out/y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/y_smoke.log 2>&1; tail -2 gpurun_out/y_smoke.log
KB2_TC_VERBOSE=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/y_bench_10m.json 2> gpurun_out/y_bench_10m.err; echo "bench exit $?"; cut -c1-3000 gpurun_out/y_bench_10m.json
KB2_PROFILE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/y_launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/y_ncu_bench.log 2>&1; echo "ncu list exit $?"
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:ivfpq_tc_filter_kernel|bound_kernel|select_keys_hist_kernel|gemm_keys_tc_kernel|finalize_warp_kernel|exact_eval_kernel' \
  -o gpurun_out/y_top_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/y_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/y_top_kernels.ncu-rep
KB2_TC_VERBOSE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/y_stage_verbose.err; grep "kb2 tc" gpurun_out/y_stage_verbose.err | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err; echo "ref exit $?"; cut -c1-700 gpurun_out/y_bench_reference.json
for w in ivf_flat_1m hnsw_1m ivf_pq_1m; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 2>gpurun_out/y_extra_$w.err >> gpurun_out/y_extra_workloads.jsonl; echo "$w exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/y_extra_workloads.jsonl'):
    if l.startswith('{'):
        j=json.loads(l); print(j['metric'], round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'), 'roofline', j['roofline'].get('bound'), round(j['roofline'].get('frac',0),3), 'cpu', (j.get('cpu_baseline') or {}).get('value'), (j.get('cpu_baseline') or {}).get('parity_vs_gpu'))
PY
