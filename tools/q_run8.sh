#!/bin/bash
mkdir -p gpurun_out
for r in 1 2; do
  NBIOT_LIB=$PWD/variants/lib_head.so timeout 200 tools/ab.sh head_$r
  timeout 200 tools/ab.sh cur_$r
  NBIOT_LIB=$PWD/variants/lib_c1.so timeout 200 tools/ab.sh c1_$r
done
grep -h "episode statistics" gpurun_out/ab_head_1.err gpurun_out/ab_cur_1.err gpurun_out/ab_c1_1.err | cut -c1-200
