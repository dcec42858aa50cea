#!/bin/bash
# QP half: GPU parity tests, then config 3 (cold + seeded map), the horizon sweep and the closed loop through bench.py
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_qp_gpu.py tests/test_roa_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --skip-e2e --skip-cpu --skip-rollout --qp-steps 5 --seed-blocks "${SEED_BLOCKS:-3x8x1x1}" > gpurun_out/bench_qp2.json 2> gpurun_out/bench_qp2.err; echo "bench exit $?"
tail -3 gpurun_out/bench_qp2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_qp2.json'))
q=d['qp']
if 'error' in q: print(q['error'])
print('qp %.4e ms %.3f frac %.3f iters %.2f'%(q['value'], q['ms_per_step'], q['roofline']['frac'], q['mean_admm_iters']))
s=q.get('seeded_map'); print('seeded %.4e ms %.3f'%(s['value'], s['ms_per_step']) if s else None)
for k,v in q['horizon_sweep'].items(): print(k, round(v['qps']), v['max_iter_count'], round(v['mean_iters'],1), '| seeded', v['seeded_map'])
print(q['closed_loop'])
PY
