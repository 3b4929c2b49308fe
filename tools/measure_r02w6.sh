#!/bin/bash
# round 2: with the class affinity in place, do the variants that lost to code size (unrolled short loops, block summaries) pay now?
for R in 1 2; do
  for V in base mlp blk; do
    unset NBIOT_LIB NBIOT_TPC_BLK
    [ $V = mlp ] && export NBIOT_LIB=$PWD/variants/lib_mlp.so
    [ $V = blk ] && export NBIOT_TPC_BLK=1
    echo -n "$V "; NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 1000000 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['N_per_gpu'], d['subframe_instances_per_s'], end='  ')
print()"
  done
done
