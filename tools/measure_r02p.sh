#!/bin/bash
# round 2: dealing cells to SMs by predicted cost (one-wave launches of the warp-per-cell kernel), on / off alternating
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dealing or builds_of_the_core or thread_per_cell or at_scale" > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02p_pytest.log
for R in 1 2 3; do
  for D in 1 0; do
    NBIOT_DEAL=$D NBIOT_BENCH_NO_CPU=1 NBIOT_BENCH_NO_BUSY=1 NBIOT_BENCH_BLOCKS=none timeout 300 python bench.py --steps 20 --warmup 3 > $O/r02p_bench_deal$D.$R.json 2> $O/r02p_bench_deal$D.$R.err
    python -c "import json; d=json.load(open('$O/r02p_bench_deal$D.$R.json')); print('deal=$D', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'faults', d['config']['faulted_instances'])"
  done
done
NBIOT_DEAL=1 python tools/inst_times.py 2>&1 | tail -4
