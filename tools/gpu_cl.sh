#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qp_gpu.py -x -q 2>&1 | tail -8
for v in 1 0; do
if [ $v = 1 ]; then export CARMPC_NO_ACTIVE_SET_REUSE=1; else unset CARMPC_NO_ACTIVE_SET_REUSE; fi
timeout 600 python bench.py --steps 20 --skip-e2e --skip-rollout --skip-cpu --skip-sweep --qp-steps 2 > gpurun_out/bench_cl_$v.json 2> gpurun_out/bench_cl_$v.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cl_$v.json'))
print('no_reuse=$v', json.dumps(d['qp']['closed_loop']))
PY
done
