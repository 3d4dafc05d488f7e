#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gemm_tc_gpu.py -m gpu -x -q -k "128-128-32 or 7-130-36" > gpurun_out/sanitizer_tc.log 2>&1
tail -15 gpurun_out/sanitizer_tc.log
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -s > gpurun_out/pytest_tc.log 2>&1
tail -30 gpurun_out/pytest_tc.log
