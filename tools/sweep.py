#!/usr/bin/env python3
"""BASELINE.json configs[4] on one GPU: config-1 traffic (M=1000, ratio=1, default agents on the device, no host
round trip), N instances advanced to a horizon in ONE launch, episode statistics reduced on the device.
usage: python tools/sweep.py [N ...]      (prints one JSON line per N)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nb_iot_environment_b200 import BatchedNBIoTEnv
CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
T = int(os.environ.get('NBIOT_SWEEP_T', '30000'))
for N in [int(x) for x in (sys.argv[1:] or ['1000', '10000', '100000'])]:
    env = BatchedNBIoTEnv(CFG1, num_envs=N, ue_cap=128, grid_w=1024, hist_cap=64)
    env.reset()
    env.run_until(3000)                      # warm-up launch
    torch.cuda.synchronize()
    s0 = env.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.run_until(T); e1.record(); torch.cuda.synchronize()
    s1 = env.stats()
    ms = e0.elapsed_time(e1)
    sf = s1['subframes'] - s0['subframes']
    print(json.dumps({'N': N, 'horizon': T, 'ms': ms, 'subframe_instances_per_s': sf / ms * 1e3, 'events_per_subframe': (s1['events'] - s0['events']) / sf,
                      'faulted': s1['faulted'], 'max_lookahead': s1['max_lookahead'], 'max_live_ues': s1['max_live_ues'],
                      'state_GB': env.state.numel() / 1e9, 'departures': sum(s1['ue_departures'])}), flush=True)
    env.close(); del env
    torch.cuda.empty_cache()
