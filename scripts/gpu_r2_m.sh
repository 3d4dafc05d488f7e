#!/bin/bash
# round 2, call M (2 GPUs): multi-GPU tests, C5 smoke at 10M rows, bench.py at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > gpurun_out/smi_m.txt; cat gpurun_out/smi_m.txt
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -x -rf > gpurun_out/pytest_m.log 2>&1; echo "exit $?" >> gpurun_out/pytest_m.log; tail -3 gpurun_out/pytest_m.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_c5.py --rows 10000000 --nlist 8192 --steps 5 --warmup 2 > gpurun_out/c5_smoke_m.json 2> gpurun_out/c5_smoke_m.err; echo "c5 smoke exit $?"; tail -c 1500 gpurun_out/c5_smoke_m.json; tail -5 gpurun_out/c5_smoke_m.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_n2_m.json 2> gpurun_out/bench_10m_n2_m.err; echo "n2 exit $?"; python -c "
import json; j=json.loads(open('gpurun_out/bench_10m_n2_m.json').read()); print('N=2 qps', round(j['value']), 'ms', round(j['ms_per_step'],3), 'recall', j['config']['recall_at_10'], j.get('multi_gpu'), j.get('multi_gpu_parity'))" || tail -5 gpurun_out/bench_10m_n2_m.err
