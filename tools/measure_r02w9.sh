#!/bin/bash
# round 2: four entries in flight per step when one lane shifts the sorted connected queue (default) against entry by entry (-DNB_CONN_SHIFT_1)
for R in 1 2; do
  for V in new old; do
    if [ $V = new ]; then unset NBIOT_LIB; else export NBIOT_LIB=$PWD/variants/lib_shift1.so; fi
    echo -n "$V loaded 65536 (tpcv forced): "; NBIOT_KERNEL=tpcv NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS=none timeout 400 python bench.py --instances 65536 --steps 6 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
    echo -n "$V sweep: "; NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['subframe_instances_per_s'])"
  done
done
