#!/bin/bash
# ADMM iteration histogram of the config-3 grid and the tail study (how much of the ADMM time a few slow samples cost)
mkdir -p gpurun_out
python tools/iter_hist.py ${HORIZON:-20} 2>&1 | tail -14
python tools/admm_tail.py 2>&1 | tail -8
