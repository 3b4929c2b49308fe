#!/bin/bash
# round 2, fourth GPU session: full GPU test suite, the new bench line, occupancy variants of the voted thread-per-cell kernel, crossover
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02d_pytest_gpu.log
timeout 900 python bench.py > $O/r02d_bench.json 2> $O/r02d_bench.err; echo "bench rc=$?"; tail -c 1500 $O/r02d_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'], 'kernel', d['config']['kernel'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'], b['max_live_ues_per_cell'], b.get('kernel'), b.get('instance_decisions_per_s'))
print('cpu', d.get('cpu_baseline'), d.get('parity_vs_reference')); print('issue', d.get('issue'))
PY
for V in F H I K J; do
  NBIOT_LIB=variants/lib_$V.so NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 1000000 > $O/r02d_sweep_$V.jsonl 2> $O/r02d_sweep_$V.err; echo "variant $V sweep rc=$?"; cut -c1-140 $O/r02d_sweep_$V.jsonl
done
for N in 16384 32768 65536; do
  NBIOT_KERNEL=warp timeout 300 python tools/sweep.py $N > $O/r02d_x_warp_$N.jsonl 2>/dev/null; echo "warp $N"; cut -c1-140 $O/r02d_x_warp_$N.jsonl
  NBIOT_LIB=variants/lib_F.so NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py $N > $O/r02d_x_F_$N.jsonl 2>/dev/null; echo "tpcv(F) $N"; cut -c1-140 $O/r02d_x_F_$N.jsonl
done
NBIOT_LIB=variants/lib_F.so NBIOT_KERNEL=tpcv timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 1 -c 1 -f -o $O/r02d_prof_tpcv_F python tools/profile_cfg5.py > $O/r02d_ncu_F.log 2>&1; echo "ncu F rc=$?"; tail -1 $O/r02d_ncu_F.log
