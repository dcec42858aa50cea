#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 3; do
CMD="python tools/admm_breakdown.py"
CARMPC_ADMM_DEBUG=$f $CMD > gpurun_out/plain_bd.log 2>&1 && \
CARMPC_ADMM_DEBUG=$f ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bd$f.csv $CMD > gpurun_out/ncu_bd.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_bd$f.csv')))
rows=[r for r in rows if len(r)>14 and r[0].isdigit()]
print('flags=$f', [(r[4].split('(')[0][-28:], round(float(r[14])/1e6,2)) for r in rows if 'admm' in r[4] or 'polish' in r[4]])
PY
done
