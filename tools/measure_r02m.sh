#!/bin/bash
# round 2, final validation: full GPU suite, smoke, the bench line and the reference arm of the final build
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02f_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err; echo "ref arm rc=$?"; cut -c1-200 $O/r02f_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err; echo "bench rc=$?"; tail -c 300 $O/r02f_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'], 'kernel', d['config']['kernel'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'], b['max_live_ues_per_cell'], b.get('instance_decisions_per_s'), b.get('issue',{}).get('inst_per_subframe_instance'))
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_port',{}).get('value'), d.get('parity_vs_reference')); print('issue', d.get('issue'))
PY
NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 1000000 > $O/r02f_sweep.jsonl 2>/dev/null; cut -c1-140 $O/r02f_sweep.jsonl
