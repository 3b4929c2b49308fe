#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r9.log
for r in 1 2; do
  NBIOT_LIB=$PWD/variants/lib_head.so timeout 200 tools/ab.sh head_$r
  timeout 200 tools/ab.sh cur_$r
  timeout 200 tools/ab.sh gen_$r NBIOT_SIM_VARIANT=generic
done
grep -h "episode statistics" gpurun_out/ab_head_1.err gpurun_out/ab_cur_1.err | cut -c1-200
