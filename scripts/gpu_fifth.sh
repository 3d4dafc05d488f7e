#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -s 2>&1 | grep -E "max err|passed|failed" > gpurun_out/pytest_tc.log
cat gpurun_out/pytest_tc.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_ivf_gpu.py -m gpu -x -q -k "imported_index_parity and 16-128 or small_batch or refine" > gpurun_out/sanitizer_ivf.log 2>&1
tail -6 gpurun_out/sanitizer_ivf.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gemm_tc_gpu.py > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
# whole suite again with the tensor-core contraction as the FLAT / coarse engine
KB2_GEMM=tc timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gemm_tc_gpu.py > gpurun_out/pytest_tcmode.log 2>&1
echo "pytest(tc mode) exit $?" >> gpurun_out/pytest_tcmode.log
tail -8 gpurun_out/pytest_tcmode.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
cat gpurun_out/bench_10m.json; tail -5 gpurun_out/bench_10m.err
KB2_GEMM=tc timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_tc.json 2> gpurun_out/bench_10m_tc.err
cat gpurun_out/bench_10m_tc.json; tail -5 gpurun_out/bench_10m_tc.err
KB2_PROFILE=1 KB2_GEMM=tc timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivfpq_scan -c 1 -o gpurun_out/prof_scan_10m -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
KB2_PROFILE=1 KB2_GEMM=tc timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_keys_tc -c 1 -o gpurun_out/prof_gemm_tc -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
tail -2 gpurun_out/ncu_full_tc.log
