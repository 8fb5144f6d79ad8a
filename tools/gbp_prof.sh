#!/bin/bash
# development aid: GNN_BP4 parity + throughput + per-kernel counters
timeout 300 python -m pytest tests/test_gnn_bp4.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/gbp_bench.py 32768
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_gbp -s 8 -c 3 --csv --log-file gpurun_out/gbp_ncu.csv python tools/gbp_bench.py 8192 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/gbp_ncu.csv')) if len(r)>10]
cur={}
for r in rows[1:]:
    k=(r[0],r[4][:30]); cur.setdefault(k,{})[r[-3].split('.')[0].replace('__','_')[:18]]=r[-1]
for k,v in cur.items(): print(k, v)
PY
