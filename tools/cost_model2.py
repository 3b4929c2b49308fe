#!/usr/bin/env python3
"""How predictable is a cell's launch cost from its LOAD STATE at launch (live UEs, list lengths, pending events) and the physical meaning
of its action (NPRACH occasions the step will hold)? Headline workload; NBIOT_PROF_INST=1. Round-2 follow-up of tools/cost_model.py."""
import os, subprocess, sys, tempfile
os.environ['NBIOT_PROF_INST'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi

# byte offsets of the header fields used as features (host compile of csrc/nbiot_state.h)
SRC = '#include <stdio.h>\n#include <stddef.h>\n#include "nbiot_state.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(NbHdr), offsetof(NbHdr, free_top), offsetof(NbHdr, conn_n), offsetof(NbHdr, cont_n), offsetof(NbHdr, ra_n), offsetof(NbHdr, pool_free), offsetof(NbHdr, n_scheduled));}\n'
d = tempfile.mkdtemp()
open(os.path.join(d, 'o.cpp'), 'w').write(SRC)
subprocess.check_call(['g++', '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(ROOT, 'nb_iot_environment_b200', 'csrc'), os.path.join(d, 'o.cpp'), '-o', os.path.join(d, 'o')])
HDR, O_FREE, O_CONN, O_CONT, O_RA, O_POOL, O_SCHED = [int(x) for x in subprocess.check_output([os.path.join(d, 'o')]).split()]

N = bench.N_PER_GPU
env = BatchedNBIoTEnv(bench.CONF, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT, num_envs=N, **bench.CAPS)
S = 12
acts = bench.ext_actions(S, N, seed=1234)
PERIODS = np.array([80, 160, 240, 320, 640, 1280, 2560])
env.reset()
X, Y, PREV = [], [], []
prev = None


def header_features():
    h = env.state[:N * HDR].cpu().numpy().reshape(N, HDR)
    i32 = lambda off, k=1: h[:, off:off + 4 * k].copy().view(np.int32).reshape(N, k)
    live = bench.CAPS['ue_cap'] - i32(O_FREE)[:, 0]
    pool = h[:, O_POOL:O_POOL + 8].copy().view(np.uint64)[:, 0]
    pending = np.array([bin(int(~p) & ((1 << 48) - 1)).count('1') for p in pool])
    return [live, i32(O_CONN)[:, 0], i32(O_CONT)[:, 0], i32(O_RA, 3).sum(1), pending, i32(O_SCHED)[:, 0]]


for s in range(S):
    f = header_features()
    env.step(torch.from_numpy(acts[s]).cuda())
    out = np.zeros((N, 4), dtype=np.uint64)
    _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    dur = (out[:, 1].astype(np.int64) - out[:, 0].astype(np.int64)) / 1e6
    if s >= 2:
        a = acts[s].astype(np.int64)                      # th_C1, th_C0, sc_C0..2, period_C0..2
        occ = (3000.0 / PERIODS[a[:, 5:8]])               # NPRACH occasions per CE level in the step (CE0 / CE1 inactive when their threshold is off)
        two_level = a[:, 0] >= a[:, 1]
        occ[:, 0] = np.where(two_level, 0.0, occ[:, 0])
        feats = [np.ones(N)] + [x.astype(np.float64) for x in f] + [occ[:, 0], occ[:, 1], occ[:, 2], occ.sum(1), (a[:, 2:5] + 1).sum(1).astype(np.float64)]
        X.append(np.stack(feats, 1)); Y.append(dur); PREV.append(prev)
    prev = dur
X, Y, PREV = np.concatenate(X), np.concatenate(Y), np.concatenate(PREV)
n = len(Y); tr = np.arange(n) % 2 == 0
names = ['1', 'live UEs', 'connected', 'in contention', 'in RA lists', 'pending events', 'scheduled', 'occ CE0', 'occ CE1', 'occ CE2', 'occ total', 'preamble groups']
w = np.linalg.lstsq(X[tr], Y[tr], rcond=None)[0]
pred = X[~tr] @ w
r2 = 1 - ((Y[~tr] - pred) ** 2).sum() / ((Y[~tr] - Y[~tr].mean()) ** 2).sum()
print('cells x steps', n, 'duration mean %.2f std %.2f ms' % (Y.mean(), Y.std()))
for j, nm in enumerate(names[1:], 1):
    print('  correlation of the duration with %-16s %.3f' % (nm, np.corrcoef(Y, X[:, j])[0, 1]))
print('  correlation of the duration with the previous step\'s duration %.3f' % np.corrcoef(Y, PREV)[0, 1])
print('R2 of a linear model on load state at launch + NPRACH occasions of the action (held-out half): %.3f' % r2)
# what dealing by predicted cost could buy: SMs hold 28 cells; if cells were dealt so that every SM gets the same PREDICTED work, the spread of
# the per-SM sums of the ACTUAL durations is what is left
k = 28
res = Y[~tr] - pred
g = len(res) // k
print('std of the per-SM sum of durations (28 cells per SM): random dealing %.2f ms, dealing by prediction leaves the residuals %.2f ms (mean sum %.1f ms)'
      % (Y[~tr][:g * k].reshape(g, k).sum(1).std(), res[:g * k].reshape(g, k).sum(1).std(), Y.mean() * k))

# the predictor a launch could afford: the cell's previous duration + two header words
for cols, label in (([0, 4], 'RA-list length'), ([0, 1, 4], 'live UEs + RA-list length')):
    for with_prev in (False, True):
        A = X[:, cols] if not with_prev else np.column_stack([X[:, cols], PREV])
        w2 = np.linalg.lstsq(A[tr], Y[tr], rcond=None)[0]
        p2 = A[~tr] @ w2
        r2b = 1 - ((Y[~tr] - p2) ** 2).sum() / ((Y[~tr] - Y[~tr].mean()) ** 2).sum()
        res2 = Y[~tr] - p2
        print('model [1, %s%s]: R2 %.3f, coefficients %s, per-SM residual std %.2f ms' % (label, ', previous duration' if with_prev else '', r2b, np.round(w2, 5).tolist(),
                                                                                       res2[:g * k].reshape(g, k).sum(1).std()))
