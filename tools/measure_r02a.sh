#!/bin/bash
# round 2, first GPU session: the thread-per-cell kernel against the warp-per-cell kernel (one GPU)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02a_pytest_gpu.log
for K in warp tpc; do
  NBIOT_KERNEL=$K timeout 300 python tools/sweep.py 100000 1000000 > $O/r02a_sweep_$K.jsonl 2> $O/r02a_sweep_$K.err; echo "sweep $K rc=$?"; cat $O/r02a_sweep_$K.jsonl
  NBIOT_KERNEL=$K NBIOT_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 8 > $O/r02a_bench_$K.json 2> $O/r02a_bench_$K.err; echo "bench $K rc=$?"; python -c "import json,sys; d=json.load(open('$O/r02a_bench_$K.json')); print(d['value'], d['ms_per_step'], d['config']['faulted_instances'], d['config']['max_live_ues_per_cell'])"
  NBIOT_KERNEL=$K NBIOT_BENCH_NO_CPU=1 timeout 300 python bench.py --steps 6 --instances 65536 > $O/r02a_bench64k_$K.json 2> $O/r02a_bench64k_$K.err; echo "bench64k $K rc=$?"; python -c "import json,sys; d=json.load(open('$O/r02a_bench64k_$K.json')); print(d['value'], d['ms_per_step'], d['config']['faulted_instances'])"
  NBIOT_KERNEL=$K timeout 300 python tools/config3_bench.py 65536 200 > $O/r02a_config3_$K.json 2> $O/r02a_config3_$K.err; echo "config3 $K rc=$?"; cat $O/r02a_config3_$K.json
done
# ncu: the thread-per-cell kernel on the sweep configuration
NBIOT_KERNEL=tpc python tools/profile_cfg5.py > $O/r02a_plain_cfg5.log 2>&1 && \
NBIOT_KERNEL=tpc ncu --set full --clock-control none --import-source on -k regex:nbiot_simt_kernel -s 1 -c 1 -f -o $O/r02a_prof_tpc python tools/profile_cfg5.py > $O/r02a_ncu_tpc.log 2>&1; echo "ncu tpc rc=$?"; tail -2 $O/r02a_ncu_tpc.log
NBIOT_KERNEL=warp ncu --set full --clock-control none --import-source on -k regex:nbiot_sim_kernel -s 1 -c 1 -f -o $O/r02a_prof_warp_cfg5 python tools/profile_cfg5.py > $O/r02a_ncu_warp.log 2>&1; echo "ncu warp rc=$?"; tail -2 $O/r02a_ncu_warp.log
