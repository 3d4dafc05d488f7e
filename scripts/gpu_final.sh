#!/bin/bash
# round-end style validation: tests, smoke, bench (ours + reference arm), ncu launch list + full captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; cat gpurun_out/bench_10m.json; tail -3 gpurun_out/bench_10m.err
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ivfpq_tc_filter|bound_kernel|lut_build|exact_eval" -c 4 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
