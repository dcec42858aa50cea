#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_qp_gpu.py tests/test_membership_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --skip-e2e --skip-cpu --qp-steps 5 > gpurun_out/bench_ffma2.json 2> gpurun_out/bench_ffma2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ffma2.json'))
print('membership %.4e ms %.4f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac']))
print('rollout %.4e ms %.4f'%(d['rollout']['value'], d['rollout']['ms_per_step']))
q=d['qp']; print('qp %.4e ms %.3f frac %.3f iters %.2f'%(q['value'], q['ms_per_step'], q['roofline']['frac'], q['mean_admm_iters']))
print({k:(round(v['qps']),v['max_iter_count'],round(v['mean_iters'],1)) for k,v in q['horizon_sweep'].items()}); print(q['closed_loop'])
PY
