#!/bin/bash
# first contact with the GPU: sanitizer on a tiny case, then the parity tests, then a throughput probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_flat_gpu.py -m gpu -x -q -k "kat or fewer or bitset" > gpurun_out/sanitizer_flat.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python scripts/quick_bench.py 1e6 10000 1024 64 > gpurun_out/quick_1m.log 2>&1
tail -5 gpurun_out/sanitizer_flat.log; tail -30 gpurun_out/pytest.log; cat gpurun_out/quick_1m.log
