#!/bin/bash
# seeded region-of-attraction map: parity tests, then the block sweep through bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qp_gpu.py -x -q -k "seeded" 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --skip-e2e --skip-cpu --skip-rollout --skip-sweep --skip-closed-loop --qp-steps 5 \
  --seed-blocks "${SEED_BLOCKS:-2x8x1x1,1x8x1x1,2x4x1x1,3x8x1x1,2x16x1x1,1x4x1x1,4x8x1x1,2x8x1x2,4x16x1x1}" > gpurun_out/bench_seeded.json 2> gpurun_out/bench_seeded.err; echo "bench exit $?"
tail -3 gpurun_out/bench_seeded.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_seeded.json'))
q=d['qp']; print('cold qp %.4e ms %.3f iters %.2f'%(q['value'], q['ms_per_step'], q['mean_admm_iters']), q.get('polish'))
if 'error' in q: print(q['error'])
for k,v in q.get('seeded_map',{}).get('by_block',{}).items(): print(k, {a:(float('%.4g'%b) if isinstance(b,float) else b) for a,b in v.items()})
PY
