#!/bin/bash
# round 2, seventh GPU session: full GPU suite on the final build, bench line, ncu launch list + full captures, compute-sanitizer
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02g_pytest_gpu.log
timeout 900 python bench.py > $O/r02g_bench.json 2> $O/r02g_bench.err; echo "bench rc=$?"; tail -c 300 $O/r02g_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'], 'kernel', d['config']['kernel'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'], b['max_live_ues_per_cell'], b.get('instance_decisions_per_s'))
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_port',{}).get('value'), d.get('parity_vs_reference')); print('issue', d.get('issue'))
PY
# launch list of the bench command (headline workload only, no CPU legs)
NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS= ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02g_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02g_ncu_l.log 2>&1; echo "ncu list rc=$?"
NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS=5 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02g_launches_cfg5.csv python bench.py --steps 2 --warmup 3 > $O/r02g_ncu_l5.log 2>&1; echo "ncu list cfg5 rc=$?"
# full captures: the headline kernel (warp-per-cell, single-carrier staged build) and the voted thread-per-cell kernel on the sweep configuration
python tools/profile_step.py > $O/r02g_plain_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbiot_sim_kernel -s 4 -c 1 -f -o $O/r02g_prof_warp python tools/profile_step.py > $O/r02g_ncu_warp.log 2>&1; echo "ncu warp rc=$?"; tail -1 $O/r02g_ncu_warp.log
python tools/profile_cfg5.py > $O/r02g_plain_cfg5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 1 -c 1 -f -o $O/r02g_prof_tpcv python tools/profile_cfg5.py > $O/r02g_ncu_tpcv.log 2>&1; echo "ncu tpcv rc=$?"; tail -1 $O/r02g_ncu_tpcv.log
# compute-sanitizer on small batches: both kernels, memcheck + racecheck (warp-per-cell uses shared memory and warp collectives)
for K in warp tpcv; do
  NBIOT_KERNEL=$K NBIOT_PROFILE_N=96 timeout 900 compute-sanitizer --tool memcheck python tools/profile_step.py > $O/r02g_memcheck_$K.log 2>&1; echo "memcheck $K rc=$?"; grep -E "ERROR SUMMARY|^ok" $O/r02g_memcheck_$K.log
done
NBIOT_KERNEL=warp NBIOT_PROFILE_N=48 timeout 900 compute-sanitizer --tool racecheck python tools/profile_step.py > $O/r02g_racecheck_warp.log 2>&1; echo "racecheck warp rc=$?"; grep -E "RACECHECK SUMMARY|^ok" $O/r02g_racecheck_warp.log
