#!/bin/bash
# instruction-supply counters of the migrating-instance kernel (one launch of the bench workload)
mkdir -p gpurun_out
M=sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
export NBIOT_SCHED=queue NBIOT_Q_HOP_MIN=64 NBIOT_Q_IDLE=2000
timeout 120 python tools/profile_step.py > gpurun_out/plain_q.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_q.log; exit 1; }
timeout 600 ncu --metrics $M --clock-control none -k regex:nbiot_simg -s 3 -c 1 --csv --log-file gpurun_out/icc_queue.csv python tools/profile_step.py > gpurun_out/ncu_q.log 2>&1
tail -3 gpurun_out/ncu_q.log
cat gpurun_out/icc_queue.csv | tail -20
