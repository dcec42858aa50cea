#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qp_gpu.py -x -q > gpurun_out/pytest_qp.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_qp.log
tail -5 gpurun_out/pytest_qp.log
for N in 20 10 40 80; do python tools/prof_qp.py $N 1000000 2 2>&1 | grep -v Using; done > gpurun_out/qp_sweep.log
cat gpurun_out/qp_sweep.log
CMD="python tools/prof_qp.py 20 1000000 2"
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_qp.csv $CMD > gpurun_out/ncu_qp1.log 2>&1
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:admm_kernel -s 2 -c 1 -o gpurun_out/prof_admm $CMD > gpurun_out/ncu_qp2.log 2>&1
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:polish_kernel -s 2 -c 1 -o gpurun_out/prof_polish $CMD > gpurun_out/ncu_qp3.log 2>&1
