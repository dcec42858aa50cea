#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_qp_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --skip-e2e --skip-cpu --skip-rollout --qp-steps 5 > gpurun_out/bench_qp2.json 2> gpurun_out/bench_qp2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_qp2.json'))
q=d['qp']; print('qp %.4e ms %.3f frac %.3f iters %.2f'%(q['value'], q['ms_per_step'], q['roofline']['frac'], q['mean_admm_iters']))
print({k:(round(v['qps']),v['max_iter_count'],round(v['mean_iters'],1)) for k,v in q['horizon_sweep'].items()}); print(q['closed_loop'])
PY
