#!/usr/bin/env python3
"""Wrapper-level fuzz against the LIVE unmodified reference: a random configuration (tests/fuzz_confs.py, one carrier), one of the
reference's gym wrappers (wrappers.py) over its own Controller + gym Environment, random wrapper-level actions; the batched wrappers
(nb_iot_environment_b200/wrappers.py) run over the CPU test double of the environment (tests/oracle_vector_env.py).
Observations and rewards are compared at every step.
usage: python tools/fuzz_ref_wrappers.py <case> [steps]      (build container only: needs /root/reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import conftest  # noqa: E402,F401
from fuzz_confs import fuzz_case  # noqa: E402
from golden_util import REF_RTOL  # noqa: E402
from test_gpu_parity import NPRACH_AGENTS, NPUSCH_AGENTS  # noqa: E402
from tools.refharness import load_reference, PhiloxShim  # noqa: E402
from oracle_vector_env import OracleVectorEnv  # noqa: E402
from nb_iot_environment_b200 import wrappers as B  # noqa: E402

k = int(sys.argv[1]); STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 0
conf, _, cap = fuzz_case(k)
r = np.random.default_rng(70_000 + k)
kind = ['nprach', 'nprach_discrete', 'nprach_default', 'scheduler', 'basic', 'rar_msg3', 'rar_departures'][k % 7]
seed = 9100 + k
m = load_reference(1)
import gymnasium as gym  # noqa: E402  (tools/ref_stubs)
import wrappers as W  # noqa: E402
import control_agents  # noqa: E402


def ref_env(agents, ext):
    node, perf, _, _ = m['create_system'](PhiloxShim(seed), dict(conf))
    if cap:
        node.m.channel.capture_effect = True
    ctrl = m['controller'].Controller(node, agents=[m['agents'].DummyAgent(dict(a)) for a in agents])
    ctrl.reset()
    ctrl.set_ext_agent(ext)
    return gym.make('gym_system:System-v1', system=ctrl)


class _Env(OracleVectorEnv):                     # the test double with the capture flag
    def __init__(self, agents, ext):
        from oracle_controller import OracleController
        super().__init__(conf, agents, ext, [seed])
        self.ctrls = [OracleController(conf, seed, agents_conf=agents, ext_agent=ext, capture_effect=cap)]


if kind.startswith('nprach'):
    conf = dict(conf)
    a_env = W.NPRACH_wrapper(ref_env(NPRACH_AGENTS, 3), discrete=kind == 'nprach_discrete', use_default_action=kind == 'nprach_default')
    b_env = B.NPRACHWrapper(_Env(NPRACH_AGENTS, 3), discrete=kind == 'nprach_discrete', use_default_action=kind == 'nprach_default')
    n_steps = STEPS or 14
    if kind == 'nprach_discrete':
        draw = lambda: int(r.integers(0, a_env.action_space.n))           # noqa: E731
    else:
        draw = lambda: [int(r.integers(0, n)) for n in a_env.action_space.nvec]     # noqa: E731
elif kind == 'scheduler':
    conf = dict(conf, mcs_automatic=False)
    a_env = W.DiscreteActions(W.BasicSchedulerWrapper(ref_env(NPUSCH_AGENTS, 0)))
    b_env = B.DiscreteActions(B.BasicSchedulerWrapper(_Env(NPUSCH_AGENTS, 0)))
    n_steps = STEPS or 400
    draw = lambda: int(r.integers(0, a_env.action_space.n))               # noqa: E731
elif kind == 'basic':
    a_env = W.BasicWrapper(ref_env(control_agents.agents_conf, 0))
    b_env = B.BasicWrapper(_Env(None, 0))
    n_steps = STEPS or 400
    draw = lambda: int(r.integers(0, 4))                                  # noqa: E731
else:
    conf = dict(conf, ce_mcs_automatic=False)
    crit = kind[4:]
    a_env = W.RAR_wrapper(ref_env(NPRACH_AGENTS, 1), reward_criteria=crit)
    b_env = B.RARWrapper(_Env(NPRACH_AGENTS, 1), reward_criteria=crit)
    n_steps = STEPS or 300
    draw = lambda: [int(r.integers(0, 3)), int(r.integers(0, 3)), int(r.integers(0, 5))]      # noqa: E731

oa, _ = a_env.reset()
ob, _ = b_env.reset()
assert np.allclose(np.asarray(oa, float).ravel(), ob.numpy()[0], rtol=REF_RTOL, atol=1e-12), 'reset'
for s in range(n_steps):
    a = draw()
    oa, ra, _, _, _ = a_env.step(a if not isinstance(a, list) else list(a))
    ob, rb, _, _, _ = b_env.step(np.asarray([a]).reshape(1, -1))
    oa = np.asarray(oa, float).ravel()
    assert np.allclose(oa, ob.numpy()[0][:len(oa)], rtol=REF_RTOL, atol=1e-12), f'case {k} {kind} step {s}: observation'
    assert abs(float(ra) - float(rb[0])) <= 1e-12 * max(1.0, abs(float(ra))), f'case {k} {kind} step {s}: reward {ra} vs {float(rb[0])}'
print('OK case', k, kind, 'steps', n_steps, conf)
