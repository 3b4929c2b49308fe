#!/bin/bash
# round 2, eleventh GPU session: loop unrolling for memory-level parallelism in the thread-per-cell kernel
mkdir -p gpurun_out
O=gpurun_out
for V in base U1 U2 base U2 U1; do
  if [ $V = base ]; then L=nb_iot_environment_b200/libnbiot_b200.so; else L=variants/lib_$V.so; fi
  NBIOT_LIB=$L NBIOT_KERNEL=tpcv timeout 300 python tools/sweep.py 131072 > $O/r02k_sweep_$V.jsonl 2> $O/r02k_sweep_$V.err; echo "variant $V sweep rc=$?"; cut -c1-140 $O/r02k_sweep_$V.jsonl
done
for V in U1 U2; do
  NBIOT_LIB=variants/lib_$V.so NBIOT_KERNEL=tpcv timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_cell" > $O/r02k_pytest_$V.log 2>&1; echo "parity $V rc=$?"; tail -1 $O/r02k_pytest_$V.log
done
