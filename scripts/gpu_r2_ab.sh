#!/bin/bash
# round 2, call AB (1 GPU): ncu --set full of the dominant kernels of C2 (IVF_FLAT list-major tcgen05) and C4 (HNSW), one launch each
mkdir -p gpurun_out
KB2_PROFILE=1 timeout 200 ncu --profile-from-start off --set full --import-source on --clock-control none --launch-count 1 \
  --kernel-name 'regex:ivfflat_tc_kernel' -o gpurun_out/ab_c2_ivfflat_tc -f python bench.py --workload ivf_flat_1m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ab_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
KB2_PROFILE=1 timeout 240 ncu --profile-from-start off --set full --import-source on --clock-control none --launch-count 1 \
  --kernel-name 'regex:hnsw_search_kernel' -o gpurun_out/ab_c4_hnsw -f python bench.py --workload hnsw_1m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ab_ncu_c4.log 2>&1; echo "ncu c4 exit $?"
ls -la gpurun_out/ab_*.ncu-rep
