#!/bin/bash
# round 2, fifth GPU session: full GPU suite, bench line, block summaries on / off on the voted thread-per-cell kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02e_pytest_gpu.log
for B in 1 0; do
  NBIOT_TPC_BLK=$B NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 1000000 > $O/r02e_sweep_blk$B.jsonl 2> $O/r02e_sweep_blk$B.err; echo "summaries=$B sweep rc=$?"; cut -c1-140 $O/r02e_sweep_blk$B.jsonl
done
NBIOT_KERNEL=tpcv NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_BLOCKS= timeout 300 python bench.py --steps 6 --instances 65536 > $O/r02e_bench64k_tpcv.json 2> $O/r02e_bench64k_tpcv.err; echo "bench64k tpcv rc=$?"; python -c "import json,sys; d=json.load(open('$O/r02e_bench64k_tpcv.json')); print(d['value'], d['ms_per_step'], d['config']['faulted_instances'])"
NBIOT_KERNEL=warp NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_BLOCKS= timeout 300 python bench.py --steps 6 --instances 65536 > $O/r02e_bench64k_warp.json 2> $O/r02e_bench64k_warp.err; echo "bench64k warp rc=$?"; python -c "import json,sys; d=json.load(open('$O/r02e_bench64k_warp.json')); print(d['value'], d['ms_per_step'], d['config']['faulted_instances'])"
timeout 900 python bench.py > $O/r02e_bench.json 2> $O/r02e_bench.err; echo "bench rc=$?"; tail -c 600 $O/r02e_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'], 'kernel', d['config']['kernel'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'], b['max_live_ues_per_cell'], b.get('kernel'), b.get('instance_decisions_per_s'))
print('cpu', d.get('cpu_baseline'), d.get('parity_vs_reference')); print('issue', d.get('issue'))
PY
NBIOT_KERNEL=tpcv timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 1 -c 1 -f -o $O/r02e_prof_tpcv python tools/profile_cfg5.py > $O/r02e_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/r02e_ncu.log
