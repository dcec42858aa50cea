#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --skip-qp --skip-cpu --skip-e2e --steps 200 --mode $MODE > gpurun_out/b_$name.json 2>> gpurun_out/bench_m.err; python - <<PY
import json
d=json.load(open('gpurun_out/b_$name.json')); print('$name', 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'value %.4e'%d['value'])
PY
}
MODE=1 run m1_tma X=1
MODE=1 run m1_plain CARMPC_NO_TMA=1
MODE=0 run m0_plain X=1
MODE=1 run m1_tma_again X=1
MODE=0 run m0_plain_again X=1
