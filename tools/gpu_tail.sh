#!/bin/bash
mkdir -p gpurun_out
python tools/iter_hist.py 2>&1 | tail -14
python tools/admm_tail.py 2>&1 | tail -8
for cap in 0 40 60 100 200; do
  CARMPC_FIRST_ITERS=$cap timeout 300 python bench.py --steps 10 --skip-e2e --skip-cpu --skip-rollout --skip-sweep --skip-closed-loop --qp-steps 5 --seed-blocks 3x8x1x1 > gpurun_out/b_cap.json 2>gpurun_out/b_cap.err
  python - <<PY
import json
q=json.load(open('gpurun_out/b_cap.json'))['qp']
print('first-pass cap $cap: cold %.3f ms iters %.2f maxiter %d | seeded %.3f ms iters %.2f flags_equal %s'%(q['ms_per_step'], q['mean_admm_iters'], q['max_iter_count'], q['seeded_map']['ms_per_step'], q['seeded_map']['by_block']['3x8x1x1']['mean_admm_iters'], q['seeded_map']['by_block']['3x8x1x1']['flags_equal_cold']))
PY
done
