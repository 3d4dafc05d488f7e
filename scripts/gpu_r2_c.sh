#!/bin/bash
# round 2, call C (2 GPUs): new-kernel tests, multi-GPU collective test, N=1 and N=2 bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/host2.txt
timeout 900 python -m pytest tests/test_ivfpq_tc_gpu.py tests/test_multigpu_gpu.py tests/test_hnsw_gpu.py -q -rf -x > gpurun_out/pytest_c1.log 2>&1; echo "exit $?" >> gpurun_out/pytest_c1.log; tail -4 gpurun_out/pytest_c1.log
timeout 900 python -m pytest tests/test_ivf_gpu.py -q -rf -x -k "typed or tc_engine or bitset_after or add_after or deterministic" > gpurun_out/pytest_c2.log 2>&1; echo "exit $?" >> gpurun_out/pytest_c2.log; tail -4 gpurun_out/pytest_c2.log
KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_c.json 2> gpurun_out/bench_10m_c.err; cut -c1-400 gpurun_out/bench_10m_c.json; grep "kb2 tc" gpurun_out/bench_10m_c.err | tail -1
KB2_TC_A_CODES=2000 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-330
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_10m_n2.json 2> gpurun_out/bench_10m_n2.err; cut -c1-2500 gpurun_out/bench_10m_n2.json; tail -5 gpurun_out/bench_10m_n2.err
