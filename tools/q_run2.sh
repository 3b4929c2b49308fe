#!/bin/bash
# sweep of the class-switch policy of the migrating-instance kernel on the headline bench
mkdir -p gpurun_out
for hm in 8 27 64; do for idle in 200 2000 20000; do
  timeout 200 tools/ab.sh q_h${hm}_i${idle} NBIOT_SCHED=queue NBIOT_Q_STATS=1 NBIOT_Q_HOP_MIN=$hm NBIOT_Q_IDLE=$idle
  grep -h "queue stats" gpurun_out/ab_q_h${hm}_i${idle}.err
done; done
