#!/bin/bash
# round-1 (session 3) measurement suite, one GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01c.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01c.log
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r01c.json
python bench.py --impl reference > gpurun_out/bench_r01c_ref.json 2> gpurun_out/bench_r01c_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_r01c_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv env NBIOT_BENCH_NO_CPU=1 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l_r01c.log 2>&1; echo "ncu list rc=$?"
python tools/profile_step.py > gpurun_out/plain_r01c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbiot_sim_kernel -s 4 -c 1 -f -o gpurun_out/prof_r01c python tools/profile_step.py > gpurun_out/ncu_f_r01c.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_f_r01c.log
M=sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:nbiot_sim_kernel -s 4 -c 1 --csv --log-file gpurun_out/icc_r01c.csv python tools/profile_step.py > /dev/null 2>&1; tail -8 gpurun_out/icc_r01c.csv | cut -d, -f13-15
