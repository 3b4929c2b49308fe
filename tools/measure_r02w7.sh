#!/bin/bash
# round 2: where the voted thread-per-cell kernel spends its instructions on LOADED cells (configs[1] scenario, 65 536 cells)
O=gpurun_out
NBIOT_KERNEL=tpcv NBIOT_PROFILE_N=65536 python tools/profile_step.py > $O/r02w7_plain.log 2>&1; tail -1 $O/r02w7_plain.log
NBIOT_KERNEL=tpcv NBIOT_PROFILE_N=65536 ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 3 -c 1 -f -o $O/r02w7_prof_tpcv_loaded python tools/profile_step.py > $O/r02w7_ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/r02w7_ncu.log
