#!/bin/bash
# development aid: parity tests + per-launch metrics + short bench on the GPU box
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_tmp.csv python tools/prof_run.py 8192 1 2>&1 | tail -1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_tmp.csv")) if len(r)>10]
half=rows[len(rows)//2+1:]
cur={}
for r in half:
    k=r[4][:34]; cur.setdefault(k,{})[r[-3].split('.')[0][:22]]=r[-1]
for k,v in cur.items(): print(k, v)
PY
python bench.py --steps 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'],'e2e', d['e2e']['value'],'bp4 launch ms', d['roofline']['launch_ms'],'frac', d['roofline']['frac'])"
