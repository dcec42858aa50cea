#!/bin/bash
# round-end style pass: GPU tests, smoke, full bench, reference arm, ncu launch list + full captures of the three kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-sweep --skip-closed-loop --skip-seeded --qp-steps 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:membership_tma -s 3 -c 1 -o gpurun_out/prof_membership_tma $CMD > gpurun_out/ncu_a.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:admm_kernel -s 2 -c 1 -o gpurun_out/prof_admm $CMD > gpurun_out/ncu_b.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:polish_kernel -s 4 -c 1 -o gpurun_out/prof_polish $CMD > gpurun_out/ncu_c.log 2>&1
bash tools/gpu_prof_seeded.sh
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print('value %.4e'%d['value'],'frac %.3f'%d['roofline']['frac'],'e2e %.3e'%d['e2e']['value'],'e2e_grid %.3e'%d['e2e_grid']['value'],'cpu %.3e'%d['cpu_baseline']['value'], d['clocks'])
q=d['qp']; print('qp %.4e'%q['value'],'ms',q['ms_per_step'],'frac',q['roofline']['frac'],'e2e %.3e'%q['e2e']['value'],'cpu',q['cpu_baseline']['value'], 'iters', q['mean_admm_iters'])
print({k:(round(v['qps']),v['max_iter_count'],round(v['mean_iters'],1), round(v['seeded_map']['qps']) if v.get('seeded_map') else None) for k,v in q['horizon_sweep'].items()}); print(q['closed_loop'])
s=q.get('seeded_map'); print('seeded map %.4e QPs/s, %.2f ms, e2e %.3e'%(s['value'], s['ms_per_step'], s['e2e']['value']) if s else None)
PY
