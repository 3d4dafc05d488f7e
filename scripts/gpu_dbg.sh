#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 128 100" "48 96 100" "16 128 2000"; do echo "== m d nq = $cfg"; timeout 60 python scripts/dbg_m48.py $cfg 2>&1 | tail -2; echo "exit $?"; done
timeout 500 python -m pytest tests/test_ivf_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_ivf.log 2>&1; echo "exit $?" >> gpurun_out/pytest_ivf.log; tail -4 gpurun_out/pytest_ivf.log
echo "== old"; KB2_LIB=knowhere_b200/lib_old.so timeout 200 python scripts/ab_scan.py 2>&1 | grep "rep1.*prefetch=1"
echo "== new"; timeout 200 python scripts/ab_scan.py 2>&1 | grep "rep1.*prefetch=1"
