#!/bin/bash
# round 2, eighth GPU session: full GPU suite; header-in-HBM and carve-out experiments on the voted thread-per-cell kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02h_pytest_gpu.log
NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02h_sweep_base.jsonl 2>/dev/null; echo "base"; cut -c1-140 $O/r02h_sweep_base.jsonl
NBIOT_LIB=variants/lib_GH.so NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02h_sweep_gh.jsonl 2> $O/r02h_sweep_gh.err; echo "header in HBM rc=$?"; cut -c1-140 $O/r02h_sweep_gh.jsonl
for C in 0 25; do
  NBIOT_TPC_CARVEOUT=$C NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02h_sweep_carve$C.jsonl 2>/dev/null; echo "carveout $C"; cut -c1-140 $O/r02h_sweep_carve$C.jsonl
done
NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02h_sweep_base2.jsonl 2>/dev/null; echo "base again"; cut -c1-140 $O/r02h_sweep_base2.jsonl
NBIOT_LIB=variants/lib_GH.so NBIOT_KERNEL=tpcv timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_cell" > $O/r02h_pytest_gh.log 2>&1; echo "GH parity rc=$?"; tail -1 $O/r02h_pytest_gh.log
