#!/bin/bash
# round 2, tenth GPU session: short launches (one external decision per launch) on the thread-per-cell kernel with the header in place in HBM
mkdir -p gpurun_out
O=gpurun_out
for V in warp tpcv_gh tpc_gh tpcv_local; do
  case $V in
    warp) E="NBIOT_KERNEL=warp";;
    tpcv_gh) E="NBIOT_KERNEL=tpcv NBIOT_LIB=variants/lib_GH.so";;
    tpc_gh) E="NBIOT_KERNEL=tpc NBIOT_LIB=variants/lib_GH.so";;
    tpcv_local) E="NBIOT_KERNEL=tpcv";;
  esac
  env $E timeout 300 python tools/config3_bench.py 65536 200 > $O/r02j_config3_$V.json 2> $O/r02j_config3_$V.err; echo "config3 $V rc=$?"; cut -c1-200 $O/r02j_config3_$V.json
done
