#!/bin/bash
mkdir -p gpurun_out
python tools/prof_seeded.py > gpurun_out/prof_seeded_plain.log 2>&1 || { tail -5 gpurun_out/prof_seeded_plain.log; exit 1; }
# the followers' polish is the 4th polish launch of a seeded solve (anchors: small cap, overflow, second-pass final)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:polish_kernel -s 3 -c 1 \
    -o gpurun_out/prof_seeded_polish python tools/prof_seeded.py > gpurun_out/ncu_sp.log 2>&1
tail -3 gpurun_out/ncu_sp.log
