#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ivfpq_tc_gpu.py -m gpu -q -x --timeout 120 -k "shards" > gpurun_out/pytest_tc.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc.log; tail -6 gpurun_out/pytest_tc.log
