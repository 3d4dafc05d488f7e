#!/bin/bash
# round 2, call P (2 GPUs): C5 at full scale (100M x 96 int8, nlist 65536, nprobe 128), lists sharded over 2 GPUs
mkdir -p gpurun_out
KB2_TC_VERBOSE=1 timeout 2000 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_c5.py --rows 100000000 --steps 10 --warmup 3 > gpurun_out/c5_full_n2.json 2> gpurun_out/c5_full_n2.err; echo "c5 exit $?"
python -c "
import json; j=json.loads([l for l in open('gpurun_out/c5_full_n2.json') if l.startswith('{')][0]); print('C5 100M N=2 qps', round(j['value']), 'ms', round(j['ms_per_step'],3), j['stage_breakdown_rank0_ms'], j['config'], j['e2e'], j['roofline'])" || tail -20 gpurun_out/c5_full_n2.err
grep "kb2 tc" gpurun_out/c5_full_n2.err | tail -2
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
