#!/bin/bash
# tensor-core ADMM kernel: comparison runs + ncu launch lists (N = 20 and 40, 10^6 states)
mkdir -p gpurun_out
for N in 20 40 10; do
  timeout 300 python tools/tc_try.py $N 1000000 2 > gpurun_out/tc_try_$N.log 2>&1; tail -12 gpurun_out/tc_try_$N.log
done
for N in 20 40; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tc_launches_$N.csv python tools/tc_try.py $N 1000000 1 > gpurun_out/tc_ncu_$N.log 2>&1
done
