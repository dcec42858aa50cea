#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
( time python bench.py ) > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
tail -4 gpurun_out/bench_full_n1.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -4 gpurun_out/bench_ref.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -6 gpurun_out/bench_n2.err
cat gpurun_out/bench_full_n1.json gpurun_out/bench_ref.json gpurun_out/bench_n2.json
