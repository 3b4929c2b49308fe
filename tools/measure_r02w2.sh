#!/bin/bash
# round 2: class affinity shared by all warps of an SM (hint in L2, -DNB_TPC_AFFINITY_SM=<bonus>) against the plain vote, sweep point, alternating
for R in 1 2; do
  for V in aff_rare3 aff_rare7 aff_rare15 aff_rare63 aff_rare255; do
    if [ $V = base ]; then unset NBIOT_LIB; else export NBIOT_LIB=$PWD/variants/lib_$V.so; fi
    echo -n "$V "; NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 1000000 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['N_per_gpu'], d['subframe_instances_per_s'], end='  ')
print()"
  done
done

