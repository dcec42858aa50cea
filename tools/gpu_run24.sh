#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_membership_gpu.py -x -q 2>&1 | tail -5
for m in 1 0; do
  timeout 300 python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 --mode $m > gpurun_out/bench_tune_m$m.json 2>> gpurun_out/bench_m.err
  CARMPC_NO_TUNE=1 timeout 300 python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 --mode $m > gpurun_out/bench_notune_m$m.json 2>> gpurun_out/bench_m.err
done
python - <<'PY'
import json
for f in ('bench_tune_m1','bench_notune_m1','bench_tune_m0','bench_notune_m0'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'],'frac %.3f'%d['roofline']['frac'], 'members', d['config']['members'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/bench_m.err
python tools/grid_time.py
