#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -30 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print('value %.4e'%d['value'],'frac %.3f'%d['roofline']['frac'],'e2e %.3e'%d['e2e']['value'],'e2e_grid %.3e'%d['e2e_grid']['value'],'cpu %.3e'%d['cpu_baseline']['value'])
q=d['qp']; print('qp %.4e'%q['value'],'ms',q['ms_per_step'],'frac',q['roofline']['frac'],'e2e %.3e'%q['e2e']['value'],'cpu',q['cpu_baseline']['value'])
print({k:(round(v['qps']),v['max_iter_count']) for k,v in q['horizon_sweep'].items()}); print(q['closed_loop'])
PY
