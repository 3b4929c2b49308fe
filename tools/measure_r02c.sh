#!/bin/bash
# round 2, third GPU session: glue cut / one-pop / occupancy variants of the voted thread-per-cell kernel (variants/lib_*.so)
mkdir -p gpurun_out
O=gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for V in A B C D E F G; do
  if [ $V = A ]; then L=nb_iot_environment_b200/libnbiot_b200.so; else L=variants/lib_$V.so; fi
  NBIOT_LIB=$L NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02c_sweep_$V.jsonl 2> $O/r02c_sweep_$V.err; echo "variant $V sweep rc=$?"; cut -c1-150 $O/r02c_sweep_$V.jsonl
  NBIOT_LIB=$L NBIOT_KERNEL=tpcv timeout 300 ncu --metrics $M --clock-control none -k regex:nbiot_simtv_kernel -s 1 -c 1 --csv --log-file $O/r02c_metrics_$V.csv python tools/profile_cfg5.py > $O/r02c_ncu_$V.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02c_metrics_$V.csv')) if len(r)>10]
print('  ', {r[-3].split('__')[-1][:38]: r[-1] for r in rows[1:]})
PY
done
for V in C G; do
  NBIOT_LIB=variants/lib_$V.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_cell" > $O/r02c_pytest_$V.log 2>&1; echo "variant $V parity rc=$?"; tail -1 $O/r02c_pytest_$V.log
done
