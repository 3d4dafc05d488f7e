#!/bin/bash
# round 2, call B: every GPU test file in its own process (a sticky CUDA error must not mask the others), smoke, benches,
# ncu launch list + full capture of the dominant kernels
mkdir -p gpurun_out
for f in tests/test_*_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -rf -x > gpurun_out/pytest_$b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_$b.log
  echo "== $b: $(tail -2 gpurun_out/pytest_$b.log | tr '\n' ' ')"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; cut -c1-1800 gpurun_out/bench_10m.json; grep "kb2 tc" gpurun_out/bench_10m.err | tail -2
timeout 600 python bench.py --workload ivf_flat_1m --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat1m.json 2> gpurun_out/bench_flat1m.err; cut -c1-1500 gpurun_out/bench_flat1m.json; tail -2 gpurun_out/bench_flat1m.err
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ivfpq_tc_filter|bound_kernel" -c 2 -o gpurun_out/prof_r2_filter -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
