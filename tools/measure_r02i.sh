#!/bin/bash
# round 2, ninth GPU session: class affinity between the warps of a CTA in the vote of the thread-per-cell kernel
mkdir -p gpurun_out
O=gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,gpu__time_duration.sum,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for V in base A16_128 A64_128 A0_1024 A16_512 A64_512 A16_1024 A64_1024 base; do
  if [ $V = base ]; then L=nb_iot_environment_b200/libnbiot_b200.so; else L=variants/lib_$V.so; fi
  NBIOT_LIB=$L NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02i_sweep_$V.jsonl 2> $O/r02i_sweep_$V.err; echo "variant $V sweep rc=$?"; cut -c1-140 $O/r02i_sweep_$V.jsonl
done
for V in A64_128 A64_1024; do
  NBIOT_LIB=variants/lib_$V.so NBIOT_KERNEL=tpcv timeout 300 ncu --metrics $M --clock-control none -k regex:nbiot_simtv_kernel -s 1 -c 1 --csv --log-file $O/r02i_metrics_$V.csv python tools/profile_cfg5.py > $O/r02i_ncu_$V.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02i_metrics_$V.csv')) if len(r)>10]
print('  $V', {r[-3].split('__')[-1][:38]: r[-1] for r in rows[1:]})
PY
done
NBIOT_LIB=variants/lib_A64_1024.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_cell" > $O/r02i_pytest.log 2>&1; echo "parity rc=$?"; tail -1 $O/r02i_pytest.log
