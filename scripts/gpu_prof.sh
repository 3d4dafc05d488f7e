#!/bin/bash
mkdir -p gpurun_out
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"bound_kernel|ivfpq_tc_filter" -c 2 -o gpurun_out/tc_10m -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc10m.log 2>&1; tail -3 gpurun_out/ncu_tc10m.log
