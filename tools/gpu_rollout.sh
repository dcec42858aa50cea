#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_membership_gpu.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 50 --skip-qp --skip-e2e > gpurun_out/bench_rollout.json 2> gpurun_out/bench_rollout.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_rollout.json'))
print('membership value %.4e frac %.3f'%(d['value'], d['roofline']['frac']))
print(json.dumps(d['rollout'], indent=1))
PY
