#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qp_gpu.py -x -q 2>&1 | tail -3
echo "== H=2 (16 warps, S=2)"; for i in 1 2; do python tools/prof_qp.py 20 1000000 3 2>&1 | grep -v Using | tail -1; done
echo "== H=1 (8 warps, S=4)"; for i in 1 2; do CARMPC_ADMM_WIDE=1 python tools/prof_qp.py 20 1000000 3 2>&1 | grep -v Using | tail -1; done
echo "== N=40 H=2"; python tools/prof_qp.py 40 1000000 3 2>&1 | grep -v Using | tail -1
echo "== N=40 H=1"; CARMPC_ADMM_WIDE=1 python tools/prof_qp.py 40 1000000 3 2>&1 | grep -v Using | tail -1
CMD="python tools/prof_qp.py 20 1000000 2"
$CMD > gpurun_out/plain_qp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_qp_h2.csv $CMD > gpurun_out/ncu_qp1.log 2>&1
