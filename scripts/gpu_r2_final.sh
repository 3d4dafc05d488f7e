#!/bin/bash
# round 2, final validation (1 GPU): every GPU test file, smoke(), default bench (with CPU baseline), reference arm,
# other workloads, ncu launch list + one --set full capture of the filter kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/final_smi.txt
for f in tests/test_*_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -x -rf > gpurun_out/final_pytest_$n.log 2>&1; echo "== $n: $(tail -1 gpurun_out/final_pytest_$n.log) exit $?"
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_10m.json 2> gpurun_out/final_bench_10m.err; echo "bench exit $?"; cat gpurun_out/final_bench_10m.json | cut -c1-2500
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref exit $?"; cat gpurun_out/final_bench_reference.json | cut -c1-900
for w in ivf_flat_1m ivf_pq_1m hnsw_100k; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/final_extra_workloads.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/final_extra_workloads.jsonl'):
    if l.startswith('{'):
        j=json.loads(l); print(j['metric'], round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'), 'roofline', j['roofline'].get('bound'), round(j['roofline'].get('frac',0),3))
PY
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none \
  --kernel-name 'regex:ivfpq_tc_filter_kernel|bound_kernel|select_keys_hist_kernel|gemm_keys_tc_kernel' \
  -o gpurun_out/final_top_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out/final_top_kernels.ncu-rep
