#!/bin/bash
# closed loop (config 4) only, three repeats
mkdir -p gpurun_out
for r in 1 2 3; do
timeout 600 python bench.py --steps 5 --skip-e2e --skip-cpu --skip-rollout --skip-sweep --skip-seeded --qp-steps 2 > gpurun_out/bench_cl.json 2> gpurun_out/bench_cl.err
python - <<'PY'
import json
q=json.load(open('gpurun_out/bench_cl.json'))['qp']
print('qp ms %.3f'%q['ms_per_step'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in q['closed_loop'].items() if k!='workload'})
PY
done
