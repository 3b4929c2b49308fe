#!/usr/bin/env python3
"""BASELINE.json configs[2] on one GPU: scripts/npusch scenario (scenarios['2000_10_B'], experiments_RL.py:40-78), the
scheduling agent (UE id, Imcs, Nrep) external with uniform-random actions, N instances (default 65536): one launch
per external decision of every instance. Reports external steps/s, simulated subframe-instances/s and the HBM bytes
the launch has to stream (header + grid of every instance in and out)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from nb_iot_environment_b200 import BatchedNBIoTEnv
from test_gpu_parity import NPUSCH_AGENTS, CFG3
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 300
env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, ue_cap=1024, grid_w=2048, hist_cap=64)
g = torch.Generator(device='cuda'); g.manual_seed(1)
nvec = torch.tensor(env.action_nvec, device='cuda')
acts = (torch.rand((STEPS + 50, N, 3), device='cuda', generator=g) * nvec).to(torch.uint8)
env.reset()
for s in range(50):
    env.step(acts[s])
torch.cuda.synchronize()
s0 = env.stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(STEPS):
    obs, rew, _, _, _ = env.step(acts[50 + s])
e1.record(); torch.cuda.synchronize()
s1 = env.stats(); ms = e0.elapsed_time(e1)
geo = env.launch_geometry()
stream_bytes = 2 * N * (geo['smem_bytes'] // geo['warps_per_block'] - 240)       # header + grid, in and out, per launch
print(json.dumps({'N': N, 'ext_steps': STEPS, 'ms_per_step': ms / STEPS, 'instance_decisions_per_s': N * STEPS / ms * 1e3,
                  'subframe_instances_per_s': (s1['subframes'] - s0['subframes']) / ms * 1e3, 'subframes_per_ext_step': (s1['subframes'] - s0['subframes']) / N / STEPS,
                  'staged_GB_per_s': stream_bytes * STEPS / ms / 1e6, 'faulted': s1['faulted'], 'max_live_ues': s1['max_live_ues'], 'max_lookahead': s1['max_lookahead'],
                  'state_GB': env.state.numel() / 1e9, 'mean_reward': float(rew.mean())}))
