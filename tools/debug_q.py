#!/usr/bin/env python3
"""Where do the states of the warp-per-instance and the migrating-instance kernels differ? (debug aid)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from test_gpu_parity import CFG3, NPUSCH_AGENTS, _random_ext_actions
from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi

def layout(N, ue_cap, grid_w, hist_cap, C, hdr):
    ra_cap = ue_cap + 64
    al = lambda x: (x + 255) & ~255
    names, o = [], 0
    for name, size in [('hdr', N * hdr), ('grid', N * C * grid_w * 2), ('ra', N * 3 * ra_cap * 16), ('conn', N * ue_cap * 16), ('cont', N * ue_cap * 16),
                       ('ue', N * ue_cap * 64), ('free', N * ue_cap * 2), ('occ', N * 312), ('incoming', N * hist_cap * 8), ('delays', N * hist_cap * 4),
                       ('service', N * hist_cap * 4), ('rec', N * 1632), ('scratch', N * ra_cap * 2)]:
        names.append((name, o, size)); o = al(o + size)
    return names

N = int(os.environ.get('DBG_N', '700')); C = int(os.environ.get('DBG_C', '2')); STEPS = int(os.environ.get('DBG_STEPS', '60')); T = int(os.environ.get('DBG_T', '6000'))
hdr = _cabi.load().nbiot_hdr_bytes()
states = []
for sched in (None, 'queue_all'):
    if sched is None: os.environ.pop('NBIOT_SCHED', None)
    else: os.environ['NBIOT_SCHED'] = sched
    env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=list(range(N)), n_carriers=C, ue_cap=512, grid_w=2048, hist_cap=64)
    env.reset()
    r = np.random.default_rng(5)
    for _ in range(STEPS):
        env.step(_random_ext_actions(r, env.action_nvec, N))
    if T > 0: env.run_until(T)
    torch.cuda.synchronize()
    st = env.state.cpu().numpy().copy()
    h = st[:N * hdr].reshape(N, hdr); h[:, hdr - 40:] = 0
    states.append(st); print(sched, env.stats()); env.close()
d = np.nonzero(states[0] != states[1])[0]
print('differing bytes', d.size, 'of', states[0].size)
for name, o, size in layout(N, 512, 2048, 64, C, hdr):
    m = d[(d >= o) & (d < o + size)]
    if m.size:
        per = size // N
        inst = (m - o) // per
        print(name, 'bytes', m.size, 'instances', np.unique(inst)[:10], 'first offsets in instance', ((m - o) % per)[:12])
# grid detail
lay = dict((n, (o, s)) for n, o, s in layout(N, 512, 2048, 64, C, hdr))
go, gs = lay['grid']
per = gs // N
g0 = states[0][go:go + gs].reshape(N, per).view(np.uint16).reshape(N, C, 2048)
g1 = states[1][go:go + gs].reshape(N, per).view(np.uint16).reshape(N, C, 2048)
h0 = states[0][:N * hdr].reshape(N, hdr)
shown = 0
for i in range(N):
    if (g0[i] != g1[i]).any() and shown < 6:
        shown += 1
        t_now, t_ahead, time = (int(h0[i, o:o + 4].view(np.int32)[0]) for o in (808, 812, 24))
        for cc in range(C):
            cols = np.nonzero(g0[i, cc] != g1[i, cc])[0]
            if cols.size:
                live = [(t_now + k) & 2047 for k in range(0, max(0, t_ahead - t_now))]
                nlive = len(set(cols.tolist()) & set(live))
                print('inst', i, 'carrier', cc, 'time', time, 't_now', t_now, 't_ahead', t_ahead, 'ring pos now', t_now & 2047, 'ahead', t_ahead & 2047,
                      'diff cols', cols.size, 'range', cols.min(), cols.max(), 'live among them', nlive, 'warp', g0[i, cc, cols[:6]].tolist(), 'queue', g1[i, cc, cols[:6]].tolist())
