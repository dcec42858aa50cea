#!/bin/bash
# QP path: parity tests, smoke, short bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qp_gpu.py -x -q > gpurun_out/pytest_qp.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_qp.log
tail -40 gpurun_out/pytest_qp.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --skip-e2e --skip-cpu > gpurun_out/bench_qp.json 2> gpurun_out/bench_qp.err; echo "bench exit $?"
tail -5 gpurun_out/bench_qp.err
cat gpurun_out/bench_qp.json
