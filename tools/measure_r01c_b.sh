#!/bin/bash
# one GPU: config-5 sweep, config 3, config 4 with the round's final build
mkdir -p gpurun_out
timeout 600 python tools/sweep.py 4096 10000 100000 1000000 2>/dev/null > gpurun_out/sweep_r01c_1gpu.jsonl; cut -c1-150 gpurun_out/sweep_r01c_1gpu.jsonl
timeout 300 python tools/config3_bench.py 2>/dev/null | tail -1 > gpurun_out/config3_r01c.json; cut -c1-250 gpurun_out/config3_r01c.json
timeout 300 python tools/config4_bench.py 2>/dev/null | tail -1 > gpurun_out/config4_r01c.json; cut -c1-250 gpurun_out/config4_r01c.json
