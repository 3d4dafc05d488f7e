#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ivf_gpu.py -m gpu -q -x -k "sharding" > gpurun_out/pytest_shard.log 2>&1
tail -30 gpurun_out/pytest_shard.log
