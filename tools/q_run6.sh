#!/bin/bash
# deep-queue regime: config-5 sweep (N instances to t = 30000 in one launch), warp-per-instance vs migrating instances
mkdir -p gpurun_out
echo warp; timeout 300 python tools/sweep.py 10000 100000 2>/dev/null | tee gpurun_out/sweep_q6_warp.jsonl | cut -c1-170
export NBIOT_SCHED=queue NBIOT_Q_HOP_MIN=64 NBIOT_Q_IDLE=2000
for b in 1 16 27; do
  echo queue batch $b; NBIOT_Q_BATCH=$b timeout 300 python tools/sweep.py 10000 100000 2>/dev/null | tee gpurun_out/sweep_q6_b$b.jsonl | cut -c1-170
done
