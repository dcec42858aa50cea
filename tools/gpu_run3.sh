#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/prof_qp.py 10 1000000 1 dump > gpurun_out/qp_n10.log 2>&1
tail -3 gpurun_out/qp_n10.log | cut -c1-3000
CMD="python tools/prof_qp.py 20 1000000 2"
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_qp.csv $CMD > gpurun_out/ncu_qp1.log 2>&1
cat gpurun_out/plain_qp.log | grep -v Using
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:admm_kernel -s 1 -c 1 -o gpurun_out/prof_admm $CMD > gpurun_out/ncu_qp2.log 2>&1
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:polish_kernel -s 1 -c 1 -o gpurun_out/prof_polish $CMD > gpurun_out/ncu_qp3.log 2>&1
ls -la gpurun_out
