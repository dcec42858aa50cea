#!/bin/bash
mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('n_gpus',d['n_gpus'],'value %.4e'%d['value'],'ms/step %.4f'%d['ms_per_step'],'frac %.3f'%d['roofline']['frac'],'e2e %.3e'%d['e2e']['value'], d['clocks'])
q=d['qp']; print('qp %.4e'%q['value'],'ms',q['ms_per_step'],'e2e %.3e'%q['e2e']['value'])
print(q.get('closed_loop'))
PY
