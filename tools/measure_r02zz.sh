#!/bin/bash
# round 2, last session: ncu launch list + full captures of the final build (header 120 B smaller, non-anchor log in the cold block, stall guard)
mkdir -p gpurun_out
O=gpurun_out
NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS=none timeout 300 python bench.py --steps 2 --warmup 3 > $O/r02zz_plain.json 2>/dev/null; echo "plain rc=$?"
NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02zz_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02zz_ncu_l.log 2>&1; echo "ncu list rc=$?"
python tools/profile_step.py > $O/r02zz_plain_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbiot_sim_kernel -s 4 -c 1 -f -o $O/r02zz_prof_warp python tools/profile_step.py > $O/r02zz_ncu_warp.log 2>&1; echo "ncu warp rc=$?"; tail -1 $O/r02zz_ncu_warp.log
python tools/profile_cfg5.py > $O/r02zz_plain_cfg5.log 2>&1 && NBIOT_KERNEL=tpcv ncu --set full --clock-control none --import-source on -k regex:nbiot_simtv_kernel -s 1 -c 1 -f -o $O/r02zz_prof_tpcv python tools/profile_cfg5.py > $O/r02zz_ncu_tpcv.log 2>&1; echo "ncu tpcv rc=$?"; tail -1 $O/r02zz_ncu_tpcv.log
