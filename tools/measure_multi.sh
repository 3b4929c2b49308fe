#!/bin/bash
# usage: tools/measure_multi.sh G -- bench.py and the config-5 sweep on G GPUs of one box
G=$1; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $G --steps 12 --warmup 3 > gpurun_out/bench_r01c_${G}gpu.json 2> gpurun_out/bench_r01c_${G}gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r01c_${G}gpu.json | cut -c1-400
timeout 900 $TR tools/sweep.py 1000 10000 100000 500000 2>/dev/null > gpurun_out/sweep_r01c_${G}gpu.jsonl; echo "sweep rc=$?"; cut -c1-160 gpurun_out/sweep_r01c_${G}gpu.jsonl
