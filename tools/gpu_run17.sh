#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qp_gpu.py -x -q 2>&1 | tail -4
for N in 20 10 40; do python tools/prof_qp.py $N 1000000 3 2>&1 | grep -v Using | tail -1; done
python tools/prof_qp.py 80 200000 3 2>&1 | grep -v Using | tail -1
python tools/iter_hist.py 20 2>&1 | grep -v Using | head -4
