#!/bin/bash
# batch-synchronous hand-out: do warps that enter a handler together run it faster?
mkdir -p gpurun_out
export NBIOT_SCHED=queue NBIOT_Q_HOP_MIN=64 NBIOT_Q_IDLE=2000 NBIOT_Q_STATS=1
for v in merged split; do
  if [ $v = split ]; then export NBIOT_LIB=$PWD/variants/lib_split.so; fi
  for b in 8 16 27; do
    timeout 200 tools/ab.sh q5_${v}_b$b NBIOT_Q_BATCH=$b
    grep -h "queue stats" gpurun_out/ab_q5_${v}_b$b.err
  done
done
