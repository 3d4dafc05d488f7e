#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "exit $?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
echo "== old"; KB2_LIB=knowhere_b200/lib_old.so timeout 200 python scripts/ab_scan.py 2>&1 | grep "rep1.*prefetch=1"
echo "== new"; timeout 200 python scripts/ab_scan.py 2>&1 | grep "rep1.*prefetch=1"
