#!/bin/bash
# round 2, twelfth GPU session: inlining policy (code size) of the thread-per-cell kernel
mkdir -p gpurun_out
O=gpurun_out
for V in base S0M3 S0M1 S1M1 S0M0 S2M4 base; do
  if [ $V = base ]; then L=nb_iot_environment_b200/libnbiot_b200.so; else L=variants/lib_$V.so; fi
  NBIOT_LIB=$L NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02l_sweep_$V.jsonl 2> $O/r02l_sweep_$V.err; echo "variant $V sweep rc=$?"; cut -c1-140 $O/r02l_sweep_$V.jsonl
done
