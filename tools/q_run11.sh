#!/bin/bash
mkdir -p gpurun_out
for r in 1 2; do
  NBIOT_LIB=$PWD/variants/lib_prev.so timeout 200 tools/ab.sh prev_$r
  timeout 200 tools/ab.sh cur_$r
done
grep -h "episode statistics" gpurun_out/ab_prev_1.err gpurun_out/ab_cur_1.err | cut -c1-200
