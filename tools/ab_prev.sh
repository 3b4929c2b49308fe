#!/bin/bash
# A/B on one box: the GPU tests, then alternating bench runs of variants/lib_prev.so (build it from the commit to compare with) and the current library
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
for r in 1 2 3; do
  NBIOT_LIB=$PWD/variants/lib_prev.so timeout 200 tools/ab.sh prev_$r
  timeout 200 tools/ab.sh cur_$r
done
grep -h "episode statistics" gpurun_out/ab_prev_1.err gpurun_out/ab_cur_1.err | cut -c1-200
