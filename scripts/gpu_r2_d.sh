#!/bin/bash
# round 2, call D (1 GPU): profiles of the v2 kernels + A/B of the phase-A table layout + HNSW / IVF_FLAT benches
mkdir -p gpurun_out
KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10m_d.json 2> gpurun_out/bench_10m_d.err; cut -c1-330 gpurun_out/bench_10m_d.json; grep "kb2 tc" gpurun_out/bench_10m_d.err | tail -1
KB2_BOUND_ROWW=32 KB2_TC_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_10m_d32.err | cut -c1-330; grep "kb2 tc" gpurun_out/bench_10m_d32.err | tail -1
KB2_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ivfpq_tc_filter|bound_kernel" -c 2 -o gpurun_out/prof_r2d_filter -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ivfflat_tc_kernel" -c 1 -o gpurun_out/prof_r2d_flat -f python bench.py --workload ivf_flat_1m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_flat.log 2>&1
timeout 900 python bench.py --workload hnsw_100k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hnsw100k.json 2> gpurun_out/bench_hnsw100k.err; cut -c1-1500 gpurun_out/bench_hnsw100k.json; tail -2 gpurun_out/bench_hnsw100k.err
