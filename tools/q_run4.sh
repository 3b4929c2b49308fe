#!/bin/bash
# merged (handler + glue) vs split (glue as its own class) task classes: throughput and SM instruction-cache hit rate
mkdir -p gpurun_out
M=sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio
export NBIOT_SCHED=queue NBIOT_Q_HOP_MIN=64 NBIOT_Q_IDLE=2000
for v in merged split; do
  if [ $v = split ]; then export NBIOT_LIB=$PWD/variants/lib_split.so; fi
  timeout 200 tools/ab.sh q4_$v NBIOT_Q_STATS=1
  grep -h "queue stats" gpurun_out/ab_q4_$v.err
  timeout 300 ncu --metrics $M --clock-control none -k regex:nbiot_simg -s 3 -c 1 --csv --log-file gpurun_out/icc_q4_$v.csv python tools/profile_step.py > gpurun_out/ncu_q4_$v.log 2>&1
  python - $v <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/icc_q4_{sys.argv[1]}.csv')) if len(r)>14 and r[0].isdigit()]
print(sys.argv[1], {r[12]: r[14] for r in rows})
PY
done
