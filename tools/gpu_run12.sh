#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qp_gpu.py -x -q 2>&1 | tail -3
python tools/prof_qp.py 20 1000000 3 2>&1 | grep -v Using | tail -1
CMD="python tools/prof_qp.py 20 1000000 2"
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_qp.csv $CMD > gpurun_out/ncu_qp1.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_qp.csv')))
rows=[r for r in rows if len(r)>14 and r[0].isdigit()]
for r in rows:
    if 'admm' in r[4] or 'polish' in r[4]:
        print(r[0], r[4].split('(')[0][-45:], r[7], r[8], float(r[14])/1e6,'ms')
PY
