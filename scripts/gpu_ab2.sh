#!/bin/bash
mkdir -p gpurun_out
echo "== old (1 chunk at a time)"; KB2_LIB=knowhere_b200/lib_old.so timeout 400 python scripts/ab_scan.py 2>&1 | grep rep1
echo "== new (2 chunks together)"; timeout 400 python scripts/ab_scan.py 2>&1 | grep rep1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
