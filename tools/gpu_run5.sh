#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_membership_gpu.py -x -q > gpurun_out/pytest_mem.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mem.log
tail -15 gpurun_out/pytest_mem.log
python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 > gpurun_out/bench_m1.json 2> gpurun_out/bench_m.err
python bench.py --skip-qp --skip-cpu --skip-e2e --steps 100 --mode 0 > gpurun_out/bench_m0.json 2>> gpurun_out/bench_m.err
cat gpurun_out/bench_m1.json gpurun_out/bench_m0.json
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-qp --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:membership_kernel -s 3 -c 1 -o gpurun_out/prof_membership2 $CMD > gpurun_out/ncu2.log 2>&1
