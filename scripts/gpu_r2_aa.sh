#!/bin/bash
# round 2, call AA (1 GPU): the final in-tree build (= the sources at HEAD): all GPU tests, smoke(), HNSW level-0 expansion with
# eight rows in flight (opt-in KB2_HNSW_KEY8=1) A/B on C4, C3 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rf > gpurun_out/aa_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/aa_pytest_gpu.log; tail -3 gpurun_out/aa_pytest_gpu.log
KB2_HNSW_KEY8=1 timeout 600 python -m pytest tests/test_hnsw_gpu.py -m gpu -q -x -rf > gpurun_out/aa_pytest_key8.log 2>&1; echo "pytest(key8) exit $?" | tee -a gpurun_out/aa_pytest_key8.log; tail -2 gpurun_out/aa_pytest_key8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/aa_smoke.log 2>&1; tail -1 gpurun_out/aa_smoke.log
for e in 0 1 0 1; do
  KB2_HNSW_KEY8=$e timeout 300 python bench.py --workload hnsw_1m --steps 10 --warmup 3 2>/dev/null | tee -a gpurun_out/aa_hnsw_ab.jsonl | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('key8=$e', round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config'].get('recall_at_10'), 'frac', round(j['roofline']['frac'],3))"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/aa_bench_10m.json | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('C3 qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), 'recall', j['config']['recall_at_10'])"
