#!/bin/bash
# ncu --set full of every polish launch of one seeded map solve (python tools/ncu_summary.py <rep> max = the followers' launch)
mkdir -p gpurun_out
python tools/prof_seeded.py > gpurun_out/prof_seeded_plain.log 2>&1 || { tail -5 gpurun_out/prof_seeded_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:polish_kernel -c 12 \
    -o gpurun_out/prof_seeded_polish -f python tools/prof_seeded.py > gpurun_out/ncu_sp.log 2>&1
tail -3 gpurun_out/ncu_sp.log
