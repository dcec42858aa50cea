#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_membership_gpu.py -x -q 2>&1 | tail -5
for m in 1 0; do
  timeout 300 python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 --mode $m > gpurun_out/bench_tma_m$m.json 2>> gpurun_out/bench_m.err
  CARMPC_NO_TMA=1 timeout 300 python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 --mode $m > gpurun_out/bench_notma_m$m.json 2>> gpurun_out/bench_m.err
done
python - <<'PY'
import json
for f in ('bench_tma_m1','bench_notma_m1','bench_tma_m0','bench_notma_m0'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, 'value %.4e'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'],'frac %.3f'%d['roofline']['frac'], d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/bench_m.err
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-qp --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:membership_tma -s 3 -c 1 -o gpurun_out/prof_membership_tma $CMD > gpurun_out/ncu2.log 2>&1
