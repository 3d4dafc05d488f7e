#!/bin/bash
# run with: gpurun --gpus 2 -- bash scripts/gpu_multi.sh 2
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
KB2_TC_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_10m_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_10m_n$N.json; grep "kb2 tc" gpurun_out/bench_n$N.err | tail -2; tail -3 gpurun_out/bench_n$N.err
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s > gpurun_out/pytest_multi.log 2>&1
tail -4 gpurun_out/pytest_multi.log
