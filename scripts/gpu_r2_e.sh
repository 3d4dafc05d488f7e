#!/bin/bash
# round 2, call E (2 GPUs): all GPU tests (per file), N=1 / N=2 bench, IVF_FLAT bench
mkdir -p gpurun_out
for f in tests/test_*_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -rf -x > gpurun_out/pytest_$b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_$b.log
  echo "== $b: $(tail -2 gpurun_out/pytest_$b.log | tr '\n' ' ')"
done
KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_e.json 2> gpurun_out/bench_10m_e.err; cut -c1-330 gpurun_out/bench_10m_e.json; grep "kb2 tc" gpurun_out/bench_10m_e.err | tail -1
timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat1m_e.json 2> gpurun_out/bench_flat1m_e.err; cut -c1-330 gpurun_out/bench_flat1m_e.json; python -c "
import json; j=json.loads(open('gpurun_out/bench_flat1m_e.json').read()); print('flat kernel_ms', j['roofline']['kernel_ms'], 'stage', j['roofline'].get('scan_stage_ms'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_10m_n2_e.json 2> gpurun_out/bench_10m_n2_e.err; cut -c1-330 gpurun_out/bench_10m_n2_e.json; python -c "
import json; j=json.loads(open('gpurun_out/bench_10m_n2_e.json').read()); print(j['multi_gpu_breakdown'], j['multi_gpu_parity'])"
