#!/bin/bash
# round 2: ncu capture of the voted thread-per-cell kernel with class affinity (final build)
O=gpurun_out
python tools/profile_cfg5.py > $O/r02w4_plain_cfg5.log 2>&1 && NBIOT_KERNEL=tpcv ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 1 -c 1 -f -o $O/r02w4_prof_tpcv python tools/profile_cfg5.py > $O/r02w4_ncu_tpcv.log 2>&1; echo "ncu tpcv rc=$?"; tail -1 $O/r02w4_ncu_tpcv.log
NBIOT_KERNEL=tpcv ncu --metrics sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:nbiot_simtv_kernel -s 1 -c 1 --csv --log-file $O/r02w4_icc.csv python tools/profile_cfg5.py > /dev/null 2>&1; echo "icc rc=$?"
