#!/bin/bash
# per-kernel durations of one seeded config-3 map solve at horizon ${HORIZON:-20} (ncu launch list; cold-cache, serialised)
mkdir -p gpurun_out
CMD="python tools/prof_seeded.py ${SEED_BLOCKS:-3x8x1x1} ${HORIZON:-20}"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -c 60 --csv --log-file gpurun_out/launches_seeded.csv $CMD > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_seeded.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
cur={}; seq=[]
for r in rows[1:]:
    key=r[iid]
    if key not in cur: cur[key]={'k':r[ik][:60]}; seq.append(key)
    cur[key][r[im]]=float(r[iv].replace(',',''))
tot=0
for k in seq:
    d=cur[k]; tot+=d.get('gpu__time_duration.sum',0)/1e3
    print(k, d['k'], '%.1f us'%(d.get('gpu__time_duration.sum',0)/1e3), '%.3g inst'%d.get('smsp__inst_executed.sum',0))
print('total %.1f us'%tot)
PY
