#!/bin/bash
# first GPU run of the migrating-instance kernel: parity through it, then A/B on the headline bench
mkdir -p gpurun_out
timeout 400 env NBIOT_SCHED=queue_all python -m pytest tests/test_gpu_parity.py tests/test_long_horizon.py -x -q -m gpu > gpurun_out/pytest_q1.log 2>&1; echo "pytest queue_all rc=$?"
tail -5 gpurun_out/pytest_q1.log
timeout 200 tools/ab.sh warp0
timeout 200 tools/ab.sh q0s NBIOT_SCHED=queue NBIOT_Q_STATS=1
timeout 200 tools/ab.sh q0 NBIOT_SCHED=queue
grep -h "episode statistics\|queue stats" gpurun_out/ab_warp0.err gpurun_out/ab_q0s.err gpurun_out/ab_q0.err
