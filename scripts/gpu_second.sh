#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
cat gpurun_out/bench_10m.json; tail -5 gpurun_out/bench_10m.err
# launch list + one full capture of the scan kernel on the 1M workload (short command under ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_1m.csv python bench.py --workload ivf_pq_1m --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 2 -c 2 -o gpurun_out/prof_scan_1m -f python bench.py --workload ivf_pq_1m --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
