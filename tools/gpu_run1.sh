#!/bin/bash
# first GPU pass: parity tests, smoke, peaks, bench (terminal-set half), ncu launch list + full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
python - > gpurun_out/peaks.log 2>&1 <<'PY'
from carmpc_b200.batch import measure_peak
import json
print(json.dumps({k: measure_peak(k) for k in ("fp32", "fp64", "hbm")}))
PY
cat gpurun_out/peaks.log
python bench.py --skip-qp > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --skip-qp --mode 0 --skip-cpu --skip-e2e > gpurun_out/bench_mode0.json 2>> gpurun_out/bench.err
cat gpurun_out/bench.json gpurun_out/bench_mode0.json
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-qp --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:membership_kernel -s 3 -c 2 -o gpurun_out/prof_membership $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
