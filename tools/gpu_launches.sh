#!/bin/bash
# per-kernel durations of one cold and one seeded config-3 solve (ncu launch list; cold-cache, serialised)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-rollout --skip-sweep --skip-closed-loop --qp-steps 1 --seed-blocks ${SEED_BLOCKS:-3x8x1x1}"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_seeded.csv $CMD > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_seeded.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
cur={}
seq=[]
for r in rows[1:]:
    key=r[iid]
    if key not in cur: cur[key]={'k':r[ik][:60]}; seq.append(key)
    cur[key][r[im]]=float(r[iv].replace(',',''))
for k in seq[-60:]:
    d=cur[k]; print(k, d['k'], '%.1f us'%(d.get('gpu__time_duration.sum',0)/1e3), '%.3g inst'%d.get('smsp__inst_executed.sum',0))
PY
