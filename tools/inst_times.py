#!/usr/bin/env python3
"""Per-instance and per-SM timing of one headline launch (NBIOT_PROF_INST=1): how long is the tail of a one-wave launch?"""
import os, sys, ctypes
os.environ['NBIOT_PROF_INST'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
N = bench.N_PER_GPU
env = BatchedNBIoTEnv(bench.CONF, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT, num_envs=N, **bench.CAPS)
acts = torch.from_numpy(bench.ext_actions(6, N, seed=1234)).cuda()
env.reset()
prev = None
for s in range(6):
    env.step(acts[s])
    out = np.zeros((N, 4), dtype=np.uint64)
    _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    t0, t1, sm = out[:, 0].astype(np.int64), out[:, 1].astype(np.int64), out[:, 2].astype(np.int64)
    start = t0.min(); dur = (t1 - t0) / 1e6; end = (t1 - start) / 1e6
    per_sm_end = np.array([end[sm == k].max() for k in np.unique(sm)])
    per_sm_n = np.array([(sm == k).sum() for k in np.unique(sm)])
    per_sm_work = np.array([dur[sm == k].sum() for k in np.unique(sm)])
    line = (f'step {s}: launch {end.max():.2f} ms | instance ms mean {dur.mean():.2f} p50 {np.median(dur):.2f} p90 {np.percentile(dur, 90):.2f} max {dur.max():.2f} '
            f'| SM end ms mean {per_sm_end.mean():.2f} min {per_sm_end.min():.2f} max {per_sm_end.max():.2f} | cells per SM {per_sm_n.min()}-{per_sm_n.max()} '
            f'| warp-busy fraction {dur.sum() / (end.max() * len(dur)):.3f}')
    if prev is not None:
        line += f' | corr with previous step {np.corrcoef(prev, dur)[0, 1]:.3f}'
    prev = dur
    print(line)
