#!/bin/bash
# round 2, call R (2 GPUs): the NCCL tests (skipped on 1-GPU boxes) and the N=2 bench line with its in-run parity proof
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r_smi.txt
timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -rf -s > gpurun_out/r_pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r_pytest_multi.log; grep -E "rank|passed|failed" gpurun_out/r_pytest_multi.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r_bench_10m_n2.json 2> gpurun_out/r_bench_n2.err; echo "bench exit $?"
cut -c1-2500 gpurun_out/r_bench_10m_n2.json; tail -3 gpurun_out/r_bench_n2.err
