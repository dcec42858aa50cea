#!/bin/bash
# ncu evidence only: launch list of a short bench run + full captures of the membership, ADMM and polish kernels (cold
# config 3) and of every polish launch of a seeded map (the numbers a run under ncu prints are never bench values)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-sweep --skip-closed-loop --skip-seeded --qp-steps 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:membership_tma -s 3 -c 1 -f -o gpurun_out/prof_membership_tma $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:admm_kernel -s 2 -c 1 -f -o gpurun_out/prof_admm $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:polish_kernel -s 4 -c 1 -f -o gpurun_out/prof_polish $CMD > gpurun_out/ncu_c.log 2>&1
bash tools/gpu_prof_seeded.sh
ls -la gpurun_out/*.ncu-rep
