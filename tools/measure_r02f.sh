#!/bin/bash
# round 2, sixth GPU session: full GPU suite, the bench line, summaries off / on (order swapped against r02e)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02f_smoke.log
for B in 0 1 0 1; do
  NBIOT_TPC_BLK=$B NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02f_sweep_blk$B.jsonl 2> $O/r02f_sweep_blk$B.err; echo "summaries=$B sweep rc=$?"; cut -c1-140 $O/r02f_sweep_blk$B.jsonl
done
timeout 900 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$?"; tail -c 400 $O/r02f_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'], 'kernel', d['config']['kernel'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'], b['max_live_ues_per_cell'], b.get('instance_decisions_per_s'))
print('cpu', d.get('cpu_baseline'), d.get('cpu_port',{}).get('value'), d.get('parity_vs_reference')); print('issue', d.get('issue'))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err; echo "ref arm rc=$?"; cut -c1-300 $O/r02f_bench_ref.json
