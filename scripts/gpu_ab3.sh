#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ivf_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_ivf.log 2>&1; echo "exit $?" >> gpurun_out/pytest_ivf.log; tail -4 gpurun_out/pytest_ivf.log
echo "== old"; KB2_LIB=knowhere_b200/lib_old.so timeout 300 python scripts/ab_scan.py 2>&1 | grep rep1
echo "== new (CTA-wide top-k, conflict-free LUT build)"; timeout 300 python scripts/ab_scan.py 2>&1 | grep rep1
