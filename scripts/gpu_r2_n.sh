#!/bin/bash
# round 2, call N (1 GPU): why is a C5-geometry search 85 ms?  launch list of the timed steps at N=1 (10M rows)
mkdir -p gpurun_out
KB2_TC_VERBOSE=1 timeout 900 python scripts/bench_c5.py --rows 10000000 --nlist 8192 --steps 3 --warmup 2 > gpurun_out/c5_smoke_n1.json 2> gpurun_out/c5_smoke_n1.err; echo "exit $?"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/c5_smoke_n1.json') if l.startswith('{')][0]); print('N=1 qps', round(j['value']), 'ms', round(j['ms_per_step'],3), j['stage_breakdown_rank0_ms'], j['config'])"; grep "kb2 tc" gpurun_out/c5_smoke_n1.err | tail -2
KB2_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_n.csv python scripts/bench_c5.py --rows 10000000 --nlist 8192 --steps 1 --warmup 2 > gpurun_out/ncu_c5_n.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c5_n.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[(r[ki][:70],r[gi])].append(float(r[vi].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:14]:
    print(f"{sum(v)/1e6:8.3f} ms  n={len(v):3d}  {k}")
PY
