#!/bin/bash
# round 2: class affinity of the voted thread-per-cell kernel adopted (hint per SM pair, bonus 1024, moved 1 time in 4): full GPU suite, A/B against -DNB_TPC_AFFINITY=0, bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 1400 python -m pytest tests -m gpu -x -q > $O/r02w3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02w3_pytest.log
for R in 1 2; do
  for V in aff noaff; do
    if [ $V = aff ]; then unset NBIOT_LIB; else export NBIOT_LIB=$PWD/variants/lib_noaff.so; fi
    echo -n "$V "; NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 32768 131072 1000000 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['N_per_gpu'], d['subframe_instances_per_s'], end='  ')
print()"
  done
done
unset NBIOT_LIB
echo -n "warp "; NBIOT_KERNEL=warp timeout 300 python tools/sweep.py 32768 2>/dev/null | cut -c1-120
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02w3_bench.json 2> $O/r02w3_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02w3_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'faults', d['config']['faulted_instances'])
for k in ('config5','config2','config3'):
    b=d.get(k)
    if b: print(k, b['value'], b['ms_per_step'], b['roofline']['frac'], b['faulted_instances'], b['truncated_lists'])
PY
