#!/bin/bash
# round 2, call A: full GPU test-suite, smoke, N=1 bench with the CPU baseline
mkdir -p gpurun_out
(nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os;print(len(os.sched_getaffinity(0)))"; free -g | head -2; nvidia-smi -L) > gpurun_out/host.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log; tail -15 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; cat gpurun_out/bench_10m.json | cut -c1-3000; tail -3 gpurun_out/bench_10m.err
