#!/bin/bash
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gemm_tc_gpu.py -m gpu -x -q -k "128-128-32 or 7-130-36" > gpurun_out/sanitizer_tc.log 2>&1
tail -12 gpurun_out/sanitizer_tc.log
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -s > gpurun_out/pytest_tc.log 2>&1
tail -25 gpurun_out/pytest_tc.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gemm_tc_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
timeout 300 python scripts/pq_quality.py 2>&1 | grep -v WARNING | tee gpurun_out/pq_quality.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
cat gpurun_out/bench_10m.json; tail -5 gpurun_out/bench_10m.err
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivfpq_scan -c 1 -o gpurun_out/prof_scan_10m -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
