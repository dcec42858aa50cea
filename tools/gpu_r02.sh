#!/bin/bash
# round-2 evidence pass: smoke, full bench, reference arm, ncu launch list, full captures of the kernels VERDICT r01 names
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep smoke
timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref exit $?"
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-sweep --skip-closed-loop --skip-seeded --qp-steps 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
OBJ=$PWD/carmpc_b200/_lib/obj
cap() {   # cap <name> <kernel regex> <launch skip> <object> <mangled substring> <command...>
  local name=$1 rx=$2 skip=$3 obj=$4 pat=$5; shift 5
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  bash tools/ncu_to_text.sh /tmp/$name.ncu-rep $OBJ/$obj $pat gpurun_out/${name}_ncu_full.txt
  rm -f /tmp/$name.ncu-rep
}
cap r02_membership_tma membership_tma 3 membership.o membership_tma_kernelILi1ELi0ELi128ELi2ELi1E $CMD
cap r02_admm_n20 admm_kernel 2 qp_admm.o admm_kernelILi2ELi2ELi1ELi2ELb1E $CMD
cap r02_polish polish_kernel 4 qp_polish.o polish_kernelILb0ELb0ELb0E $CMD
python tools/prof_rollout.py 4 > gpurun_out/plain_r.log 2>&1 && \
cap r02_rollout_screen membership_tma 3 membership.o membership_tma_kernelILi1ELi1ELi128ELi2ELi1E python tools/prof_rollout.py 6
python tools/prof_qp.py 40 300000 1 > gpurun_out/plain_q.log 2>&1 && \
cap r02_admm_tc_n40 admm_tc 0 qp_admm_tc.o admm_tc_kernelILi80E python tools/prof_qp.py 40 300000 1
cap r02_admm_n80 admm_kernel 0 qp_admm.o admm_kernelILi1ELi1ELi4ELi7ELb0E python tools/prof_qp.py 80 150000 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_qp40_launches.csv python tools/prof_qp.py 40 1000000 1 > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/r02_*_ncu_full.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print('value %.4e'%d['value'],'frac %.3f'%d['roofline']['frac'],'e2e %.3e'%d['e2e']['value'],'e2e_grid %.3e'%d['e2e_grid']['value'],'cpu %.3e'%d['cpu_baseline']['value'], d['clocks'])
print(d.get('qp_summary'))
q=d['qp']
print({k:(round(v['qps']),v['admm_kernel'][:8],v.get('ffma_kernel_ms')) for k,v in q['horizon_sweep'].items()}); print(q['closed_loop'])
PY
