#!/usr/bin/env python3
"""Small fixed command for ncu on the sweep configuration (BASELINE.json configs[4]: config-1 traffic, default agents on the
device): N cells (default 131072), reset, three launches of 3000 subframes each (sim-kernel launches #2..#4 after construct + reset)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nb_iot_environment_b200 import BatchedNBIoTEnv  # noqa: E402

CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
N = int(os.environ.get('NBIOT_PROFILE_N', '131072'))
env = BatchedNBIoTEnv(CFG1, num_envs=N, ue_cap=256, grid_w=1024, hist_cap=64)
env.reset()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for k in range(3):
    e[k].record()
    env.run_until(3000 * (k + 1))
e[3].record()
torch.cuda.synchronize()
st = env.stats()
print('ok', st['subframes'], st['events'], st['faulted'], [round(e[k].elapsed_time(e[k + 1]), 3) for k in range(3)])
