#!/usr/bin/env python3
"""Golden traces of the reference's gym WRAPPERS (reference wrappers.py) on top of its Controller + gym Environment,
driven by random wrapper-level actions: tests/golden/wrappers_*.npz.  Build container only (needs /root/reference).

  nprach : scripts/nprach set-up (evaluate_RL.py:46-131): NPRACH_wrapper over agent 3, MultiDiscrete([105,4,4,4,6,6,6])
  npusch : scripts/npusch set-up (experiments_RL.py:40-131): BasicSchedulerWrapper + DiscreteActions over agent 0
  basic  : BasicWrapper (wrappers.py:47-82) over agent 0 of the reference's six default agents (control_agents.py:29-71: UE selection only)
  rar    : RAR_wrapper (wrappers.py:366-421) over agent 1 of the scripts/nprach agent set, reward criteria 'msg3' and 'departures'
"""
import json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from tools.refharness import load_reference, PhiloxShim   # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_parity import NPRACH_AGENTS, NPUSCH_AGENTS, CFG1, CFG3   # noqa: E402  (agent dictionaries of the two scripts)


TRACE_METRICS = ['departures', 'NPRACH_occupation', 'service_times', 'delays', 'incoming', 'detection_ratios', 'av_delay']   # evaluate_RL.py:37 + list-valued ones


def make_env(m, conf, agents, ext, seed):
    import gymnasium as gym
    node, perf, _, _ = m['create_system'](PhiloxShim(seed), dict(conf))
    ag = [m['agents'].DummyAgent(dict(a)) for a in agents]
    ctrl = m['controller'].Controller(node, agents=ag)
    ctrl.reset()
    ctrl.set_ext_agent(ext)
    return gym.make('gym_system:System-v1', system=ctrl)


def main():
    m = load_reference(1)
    import wrappers as W
    out = os.path.join(ROOT, 'tests', 'golden')
    # ---- NPRACH wrapper
    conf = dict(CFG1, ratio=0.5)
    res = {}
    for si, seed in enumerate([500, 501]):
        # the subclass evaluate_RL.py:141 uses: same observations and rewards as NPRACH_wrapper, plus one sample per metric and step
        env = W.NPRACH_wrapper_traces(make_env(m, conf, NPRACH_AGENTS, 3, seed), TRACE_METRICS)
        rng = np.random.default_rng(seed)
        obs0, _ = env.reset()
        acts, obs, rew = [], [np.asarray(obs0, float)], []
        for _ in range(16):
            a = [int(rng.integers(0, n)) for n in env.action_space.nvec]
            o, r, _, _, info = env.step(list(a))
            acts.append(a); obs.append(np.asarray(o, float)); rew.append(float(r))
        res[f'acts_{si}'] = np.asarray(acts, np.int64); res[f'obs_{si}'] = np.asarray(obs); res[f'rew_{si}'] = np.asarray(rew)
        for metric in TRACE_METRICS:
            res[f'samples_{metric}_{si}'] = np.asarray(env.samples[metric], dtype=np.float64)
    np.savez_compressed(os.path.join(out, 'wrappers_nprach.npz'),
                        meta=np.frombuffer(json.dumps({'conf': conf, 'seeds': [500, 501], 'metrics': TRACE_METRICS}).encode(), dtype=np.uint8), **res)
    # ---- scheduler wrapper + discrete actions
    res = {}
    for si, seed in enumerate([510, 511]):
        env = W.DiscreteActions(W.BasicSchedulerWrapper(make_env(m, CFG3, NPUSCH_AGENTS, 0, seed)))
        rng = np.random.default_rng(seed)
        obs0, _ = env.reset()
        acts, obs, rew = [], [np.asarray(obs0, float)], []
        for _ in range(500):
            a = int(rng.integers(0, env.action_space.n))
            o, r, _, _, info = env.step(a)
            acts.append(a); obs.append(np.asarray(o, float)); rew.append(float(r))
        res[f'acts_{si}'] = np.asarray(acts, np.int64); res[f'obs_{si}'] = np.asarray(obs); res[f'rew_{si}'] = np.asarray(rew)
    np.savez_compressed(os.path.join(out, 'wrappers_npusch.npz'), meta=np.frombuffer(json.dumps({'conf': CFG3, 'seeds': [510, 511]}).encode(), dtype=np.uint8), **res)
    # ---- BasicWrapper: the agent selects the UE only (scalar action); default agents of control_agents.py
    import control_agents
    res = {}
    for si, seed in enumerate([530, 531]):
        env = W.BasicWrapper(make_env(m, CFG1, control_agents.agents_conf, 0, seed))
        rng = np.random.default_rng(seed)
        obs0, _ = env.reset()
        acts, obs, rew = [], [np.asarray(obs0, float)], []
        for _ in range(500):
            a = int(rng.integers(0, 4))
            o, r, _, _, info = env.step(a)
            acts.append(a); obs.append(np.asarray(o, float)); rew.append(float(r))
        res[f'acts_{si}'] = np.asarray(acts, np.int64); res[f'obs_{si}'] = np.asarray(obs); res[f'rew_{si}'] = np.asarray(rew)
    np.savez_compressed(os.path.join(out, 'wrappers_basic.npz'), meta=np.frombuffer(json.dumps({'conf': CFG1, 'seeds': [530, 531]}).encode(), dtype=np.uint8), **res)
    # ---- RAR wrapper (reference wrappers.py:366-421) over agent 1 (RAR_window: ce_level, rar_Imcs, Nrep), both reward criteria
    conf = dict(CFG1, ratio=0.5, ce_mcs_automatic=False)
    res = {}
    seeds, criteria = [520, 521], ['msg3', 'departures']
    for si, (seed, crit) in enumerate(zip(seeds, criteria)):
        env = W.RAR_wrapper(make_env(m, conf, NPRACH_AGENTS, 1, seed), reward_criteria=crit)
        rng = np.random.default_rng(seed)
        obs0, _ = env.reset()
        acts, obs, rew = [], [np.asarray(obs0, float)], []
        for _ in range(400):
            a = [int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 5))]
            o, r, _, _, info = env.step(list(a))
            acts.append(a); obs.append(np.asarray(o, float)); rew.append(float(r))
        res[f'acts_{si}'] = np.asarray(acts, np.int64); res[f'obs_{si}'] = np.asarray(obs).reshape(len(obs), -1); res[f'rew_{si}'] = np.asarray(rew)
    np.savez_compressed(os.path.join(out, 'wrappers_rar.npz'),
                        meta=np.frombuffer(json.dumps({'conf': conf, 'seeds': seeds, 'criteria': criteria}).encode(), dtype=np.uint8), **res)
    print('wrote wrapper goldens')


if __name__ == '__main__':
    main()
