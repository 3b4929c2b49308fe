#!/bin/bash
# round 2: 16-byte loads in the scalar column scan of the thread-per-cell kernel (default) against the 8-byte loop (-DNB_SCAN_8B), sweep point, alternating
for R in 1 2 3; do
  for V in new old; do
    if [ $V = new ]; then unset NBIOT_LIB; else export NBIOT_LIB=$PWD/variants/lib_scan8.so; fi
    echo -n "$V "; NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['subframe_instances_per_s'], d['ms'])"
  done
done
