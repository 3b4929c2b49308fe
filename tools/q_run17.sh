#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r17.log
for q in 0 1000 500 250 0 500; do
  timeout 200 tools/ab.sh q$q NBIOT_QUANTUM=$q
done
grep -h "episode statistics" gpurun_out/ab_q0.err gpurun_out/ab_q250.err | cut -c1-220
