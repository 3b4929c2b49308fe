#!/usr/bin/env python3
"""How predictable is a cell's launch cost from its action and state? (headline workload; NBIOT_PROF_INST=1)"""
import os, sys
os.environ['NBIOT_PROF_INST'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
N = bench.N_PER_GPU
env = BatchedNBIoTEnv(bench.CONF, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT, num_envs=N, **bench.CAPS)
S = 10
acts = bench.ext_actions(S, N, seed=1234)
env.reset()
X, Y, EV = [], [], []
prev_events = None
for s in range(S):
    recs_before = env.records()
    hdr_events = env.stats()['events']
    env.step(torch.from_numpy(acts[s]).cuda())
    out = np.zeros((N, 4), dtype=np.uint64)
    _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    dur = (out[:, 1].astype(np.int64) - out[:, 0].astype(np.int64)) / 1e6
    recs = env.records()
    ev = recs['events'].astype(np.int64)
    d_ev = ev - (prev_events if prev_events is not None else 0)
    prev_events = ev
    if s >= 2:
        a = acts[s].astype(np.int64)
        feats = [np.ones(N)]
        nv = env.action_nvec
        for j in range(a.shape[1]):
            for v in range(int(nv[j])):
                feats.append((a[:, j] == v).astype(np.float64))
        feats.append(recs_before['total_ues'].astype(np.float64))
        X.append(np.stack(feats, 1)); Y.append(dur); EV.append(d_ev)
X, Y, EV = np.concatenate(X), np.concatenate(Y), np.concatenate(EV)
n = len(Y); tr = np.arange(n) % 2 == 0
w = np.linalg.lstsq(X[tr].T @ X[tr] + 1e-3 * np.eye(X.shape[1]), X[tr].T @ Y[tr], rcond=None)[0]
pred = X[~tr] @ w
r2 = 1 - ((Y[~tr] - pred) ** 2).sum() / ((Y[~tr] - Y[~tr].mean()) ** 2).sum()
print('cells x steps', n, 'duration mean %.2f std %.2f ms' % (Y.mean(), Y.std()))
print('R2 of duration from one-hot action + connected UEs before the step (held-out half): %.3f' % r2)
print('correlation duration vs events processed in the step: %.3f' % np.corrcoef(Y, EV)[0, 1])
