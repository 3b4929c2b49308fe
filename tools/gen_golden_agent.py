#!/usr/bin/env python3
"""Golden traces of the reference's model-based NPRACH agent (agent_nprach.NPRACH_THAgent) and of the wrapper
that drives it (wrappers.NPRACH_agent_wrapper): tests/golden/mb_agent.npz.  Build container only.

  loop_<k>  : scripts/nprach/evaluate_MBRL.py set-up (:52-160) -- reference environment + Controller (agent 3 external)
              + NPRACH_agent_wrapper(n_actions=26, bounds=[0.3, 2.0], obs_items=[0,1,2,3,4,5,9,10,11,12]) driven by
              seeded random margin indices; per step: the agent's inputs (from info), its 8-item action, its arrival
              estimate, the wrapper's observation and reward.
  synth_<k> : the agent alone on synthetic info dicts (loads far above what the cell produces) so that the
              arrival_rate > max_value branch -- the only one that changes the agent's own (nsc, period) memory and with
              it the N = 12 row of the ML estimator -- and the no_th variant are exercised.
"""
import json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from tools.refharness import load_reference, PhiloxShim   # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_parity import NPRACH_AGENTS, CFG1   # noqa: E402

MAX_OCC = 48
METRICS = ['departures', 'NPRACH_occupation', 'service_times', 'beta']


def pack_inputs(info):
    det = np.zeros((3, MAX_OCC), np.int16); col = np.zeros((3, MAX_OCC), np.int16); nocc = np.zeros(3, np.int32)
    for ce in range(3):
        d, c = info['NPRACH_detection'][ce], info['NPRACH_collision'][ce]
        assert len(d) == len(c) <= MAX_OCC
        nocc[ce] = len(d); det[ce, :len(d)] = d; col[ce, :len(c)] = c
    return det, col, nocc, np.asarray(info['distribution'], float), len(info['incoming'])


class Rec:
    def __init__(self):
        self.rows = {k: [] for k in ('det', 'col', 'nocc', 'dist', 'ninc', 'margin', 'out', 'lam', 'k', 'm3', 'obs', 'rew', 'act',
                                     's_departures', 's_occupation', 's_service', 's_beta')}

    def add(self, packed, agent, out, margin, obs=None, rew=0.0, act=0):
        det, col, nocc, dist, ninc = packed
        r = self.rows
        r['det'].append(det); r['col'].append(col); r['nocc'].append(nocc); r['dist'].append(dist); r['ninc'].append(ninc)
        r['margin'].append(margin); r['out'].append(np.asarray(out, np.int64)); r['lam'].append(agent.lambda_estimate); r['k'].append(agent.k)
        r['m3'].append(np.asarray(agent.msg3_detection, float)); r['rew'].append(rew); r['act'].append(act)
        if obs is not None:
            r['obs'].append(np.asarray(obs, float))
        s = agent.samples
        r['s_departures'].append(s['departures'][-1]); r['s_occupation'].append(s['NPRACH_occupation'][-1])
        r['s_service'].append(s['service_times'][-1]); r['s_beta'].append(s['beta'][-1])

    def arrays(self, prefix):
        return {f'{prefix}_{k}': np.asarray(v) for k, v in self.rows.items() if len(v)}


def main():
    m = load_reference(1)
    cwd = os.getcwd(); os.chdir(os.environ.get('NBIOT_REFERENCE', '/root/reference'))
    import agent_nprach as A
    import wrappers as W
    import gymnasium as gym
    os.chdir(cwd)
    res, meta = {}, {'loops': [], 'synth': []}
    # ---- closed loop on the reference environment
    cases = [({'ratio': 0.5}, 700, False), ({'ratio': 0.5, 'random': True}, 701, False), ({'ratio': 1.0}, 702, True)]
    for ci, (extra, seed, no_th) in enumerate(cases):
        conf = dict(CFG1, **extra)
        node, perf, _, _ = m['create_system'](PhiloxShim(seed), dict(conf))
        ag = [m['agents'].DummyAgent(dict(a)) for a in NPRACH_AGENTS]
        ctrl = m['controller'].Controller(node, agents=ag)
        ctrl.reset(); ctrl.set_ext_agent(3)
        agent = A.NPRACH_THAgent(dict(NPRACH_AGENTS[3]), METRICS, no_th=no_th)
        env = W.NPRACH_agent_wrapper(gym.make('gym_system:System-v1', system=ctrl), agent, n_actions=26, obs_items=[0, 1, 2, 3, 4, 5, 9, 10, 11, 12],
                                     obs_window=1, bounds=[0.3, 2.0], discrete=True, norm=100)
        rng = np.random.default_rng(seed)
        obs0, info0 = env.reset()
        rec = Rec()
        steps = 40
        first_obs = np.asarray(obs0, float)
        for _ in range(steps):
            a = int(rng.integers(0, 26))
            info_in = pack_inputs(env.info)         # what the agent is about to look at (copied: the cell updates some lists in place)
            o, r, _, _, _ = env.step(a)
            rec.add(info_in, agent, env.conf, float(env.action_values[a]), obs=o, rew=r, act=a)
        res.update(rec.arrays(f'loop_{ci}')); res[f'loop_{ci}_obs0'] = first_obs
        meta['loops'].append({'conf': conf, 'seed': seed, 'no_th': no_th, 'steps': steps})
    # ---- the agent alone on synthetic info dicts
    for si, (seed, no_th, load) in enumerate([(900, False, 30), (901, False, 4), (902, True, 12), (903, False, 60)]):
        rng = np.random.default_rng(seed)
        agent = A.NPRACH_THAgent(dict(NPRACH_AGENTS[3]), METRICS, no_th=no_th)
        rec = Rec()
        for step in range(40):
            info = {'NPRACH_detection': [], 'NPRACH_collision': [], 'departures': int(rng.integers(0, 50)), 'NPRACH_occupation': float(rng.random()),
                    'service_times': [int(x) for x in rng.integers(100, 5000, size=int(rng.integers(0, 6)))]}
            for ce in range(3):
                n = int(rng.integers(0, MAX_OCC + 1)) if step % 7 else 0
                info['NPRACH_detection'].append([int(x) for x in rng.integers(0, load + 1, size=n)])
                info['NPRACH_collision'].append([int(x) for x in rng.integers(0, 1 + load // 3, size=n)])
            dist = rng.random(15); dist[rng.integers(0, 15, size=4)] = 0.0; dist = dist / dist.sum()
            info['distribution'] = [float(x) for x in dist]
            info['incoming'] = [0.0] * int(rng.integers(0, 40))
            margin = float(rng.uniform(0.3, 3.4))
            agent.set_parameter(margin)
            out = agent.get_action(None, 0.0, info, None)
            rec.add(pack_inputs(info), agent, out, margin)
        res.update(rec.arrays(f'synth_{si}'))
        meta['synth'].append({'seed': seed, 'no_th': no_th, 'load': load})
    out = os.path.join(ROOT, 'tests', 'golden', 'mb_agent.npz')
    np.savez_compressed(out, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **res)
    print('wrote', out, {k: v.shape for k, v in res.items() if k.endswith('_out')})


if __name__ == '__main__':
    main()
