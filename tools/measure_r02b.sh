#!/bin/bash
# round 2, second GPU session: voted dispatch on the thread-per-cell kernel; A/B of the all-combination edge test (warp kernel)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_cell or golden or controller_split or run_until" > $O/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02b_pytest_gpu.log
for K in tpcv tpc; do
  NBIOT_KERNEL=$K timeout 300 python tools/sweep.py 131072 1000000 > $O/r02b_sweep_$K.jsonl 2> $O/r02b_sweep_$K.err; echo "sweep $K rc=$?"; cat $O/r02b_sweep_$K.jsonl
done
NBIOT_KERNEL=tpcv NBIOT_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 6 --instances 65536 > $O/r02b_bench64k_tpcv.json 2> $O/r02b_bench64k_tpcv.err; echo "bench64k tpcv rc=$?"; python -c "import json,sys; d=json.load(open('$O/r02b_bench64k_tpcv.json')); print(d['value'], d['ms_per_step'], d['config']['faulted_instances'])"
NBIOT_KERNEL=tpcv timeout 300 python tools/config3_bench.py 65536 200 > $O/r02b_config3_tpcv.json 2> $O/r02b_config3_tpcv.err; echo "config3 tpcv rc=$?"; cat $O/r02b_config3_tpcv.json
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
NBIOT_KERNEL=tpcv ncu --metrics $M --clock-control none -k regex:nbiot_simtv_kernel -s 1 -c 1 --csv --log-file $O/r02b_tpcv_metrics.csv python tools/profile_cfg5.py > $O/r02b_ncu_tpcv.log 2>&1; echo "ncu tpcv rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02b_tpcv_metrics.csv')) if len(r)>10]
for r in rows[1:]:
    print(r[-3], r[-2], r[-1])
PY
# warp kernel, bench workload: the all-combination edge test of the first-fit search on / off, alternating on the same box
for R in 1 2; do
  NBIOT_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 8 > $O/r02b_fit_on_$R.json 2>/dev/null; python -c "import json; d=json.load(open('$O/r02b_fit_on_$R.json')); print('fast fit on ', d['value'], d['ms_per_step'])"
  NBIOT_LIB=variants/lib_nofit.so NBIOT_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 8 > $O/r02b_fit_off_$R.json 2>/dev/null; python -c "import json; d=json.load(open('$O/r02b_fit_off_$R.json')); print('fast fit off', d['value'], d['ms_per_step'])"
done
