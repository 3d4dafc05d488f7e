#!/bin/bash
# round 2, call Z (2 GPUs): NCCL tests, C5 at full scale (100M x 96 int8, m48, nlist 65536, nprobe 128) sharded over 2 GPUs,
# C3 bench at N=2 with the in-run parity proof
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/z_smi.txt
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -rf -s > gpurun_out/z_pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/z_pytest_multi.log; grep -E "rank|passed|failed" gpurun_out/z_pytest_multi.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/z_bench_10m_n2.json 2> gpurun_out/z_bench_n2.err; echo "bench exit $?"
cut -c1-1800 gpurun_out/z_bench_10m_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_c5.py --rows 100000000 --steps 10 --warmup 3 > gpurun_out/z_c5_full_n2.json 2> gpurun_out/z_c5_full_n2.err; echo "c5 exit $?"
cut -c1-2500 gpurun_out/z_c5_full_n2.json; tail -3 gpurun_out/z_c5_full_n2.err
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
