"""Fuzz: random configurations (tests/fuzz_confs.py) under a random NodeB-level policy.
 - the C oracle against the LIVE unmodified reference, decision by decision (build container only: needs /root/reference)
 - the kernel source on the CPU (1 lane and 32 emulated lanes) against the oracle
 - the CUDA kernels through the C ABI against the oracle (-m gpu)
None of the cases is in a golden fixture."""
import os

import numpy as np
import pytest

from golden_util import record_diff, REF_RTOL
from fuzz_confs import fuzz_case, random_action

REFERENCE = os.environ.get('NBIOT_REFERENCE', '/root/reference')
CAPS = dict(grid_w=4096, hist_cap=1024)


def _drive(make_a, make_b, k, seeds, decisions, rtol, what):
    conf, nc, cap = fuzz_case(k)
    a, b = make_a(conf, nc, cap, seeds), make_b(conf, nc, cap, seeds)
    ra, rb = a.reset(), b.reset()
    prng = np.random.default_rng(100 + k)
    for d in range(decisions + 1):
        for i in range(len(seeds)):
            diff = record_diff(ra[i], rb[i], rtol=rtol)
            assert not diff, f'{what} case {k} {conf} carriers {nc} capture {cap} seed {seeds[i]} decision {d}: {diff[:3]}'
        assert not (np.asarray([r['fault'] for r in rb]) & 0x1F).any(), f'{what} case {k}: fault'
        acts = np.asarray([random_action(int(ra[i]['state']), prng, nc) for i in range(len(seeds))], dtype=np.uint8)
        ra, rb = a.step(acts), b.step(acts)
    if hasattr(a, 'extras') and hasattr(b, 'extras'):       # the sample lists (traces / statistics) and the per-departure histories since reset
        for i, (xa, xb) in enumerate(zip(a.extras(conf), b.extras(conf))):
            assert set(xa) == set(xb)
            for key in xa:
                assert np.array_equal(xa[key], xb[key]), f'{what} case {k} seed {seeds[i]}: {key} differs'
    return a, b


class _Oracles:
    def __init__(self, conf, nc, cap, seeds):
        from oracle.oracle_py import OracleCell
        self.cells = [OracleCell(conf, s, n_carriers=nc, capture_effect=cap) for s in seeds]

    def reset(self):
        return [c.reset() for c in self.cells]

    def step(self, acts):
        return [c.step(a) for c, a in zip(self.cells, acts)]

    def extras(self, conf):
        out = []
        for c in self.cells:
            x = {}
            if conf.get('traces') or conf.get('statistics'):
                x['samples'] = c.samples()
            if conf.get('statistics'):
                st = c.stats()
                x.update({'stat_' + n: st[n] for n in ('delay', 'connection', 'access', 'bits')})
            out.append(x)
        return out


class _Emu:
    def __init__(self, lanes):
        self.lanes = lanes

    def __call__(self, conf, nc, cap, seeds):
        from hostemu_py import EmuBatch
        self.b = EmuBatch(conf, seeds, n_carriers=nc, lanes=self.lanes, capture_effect=cap, ue_cap=conf['M'], **CAPS)
        return self

    def reset(self):
        return self.b.reset()

    def step(self, acts):
        recs, steps = self.b.advance(acts, max_steps=1)
        assert (steps == 1).all()
        return recs

    def extras(self, conf):
        out = []
        for i in range(self.b.n):
            x = {}
            if conf.get('traces') or conf.get('statistics'):
                rec, n = self.b.samples(i)
                assert n == len(rec)
                x['samples'] = rec
            if conf.get('statistics'):
                st = self.b.stats(i)
                x.update({'stat_' + n: st[n] for n in ('delay', 'connection', 'access', 'bits')})
            out.append(x)
        return out


class _Cuda:
    def __call__(self, conf, nc, cap, seeds):
        from nb_iot_environment_b200 import BatchedNBIoTEnv
        self.env = BatchedNBIoTEnv(conf, num_envs=len(seeds), seeds=seeds, n_carriers=nc, all_states_external=True, capture_effect=cap, ue_cap=conf['M'], **CAPS)
        return self

    def reset(self):
        self.env.reset()
        return self.env.records()

    def step(self, acts):
        self.env.node_step(acts)
        return self.env.records()

    def extras(self, conf):
        out = [{} for _ in range(self.env.num_envs)]
        if conf.get('traces') or conf.get('statistics'):
            for x, (rec, n) in zip(out, self.env.samples()):
                assert n == len(rec)
                x['samples'] = rec
        if conf.get('statistics'):
            N, cap = self.env.num_envs, self.env.stat_cap
            cnt = self.env.stat_counts().cpu().numpy()
            raw = self.env.state[self.env._stat_off:self.env._stat_off + N * 4 * cap * 4].cpu().numpy().view(np.int32).reshape(N, 4, cap)
            for i, x in enumerate(out):
                assert cnt[i] <= cap
                x.update({'stat_' + n: raw[i, j, :cnt[i]].copy() for j, n in enumerate(('delay', 'connection', 'access', 'bits'))})
        return out


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the live reference is only present in the build container')
@pytest.mark.parametrize('k', range(10))
def test_oracle_against_the_live_reference(k, oracle_lib):
    """in a process of its own: the reference fixes its carrier count at import time and puts top-level modules on sys.path"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'fuzz_ref_oracle.py'), str(k), '500'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the live reference is only present in the build container')
@pytest.mark.parametrize('k', [0, 3, 8, 14, 21, 33])
def test_controller_split_against_the_live_reference(k, oracle_lib):
    """Controller level: the reference's Controller + DummyAgents (random agent set, external agent, fixed and external actions, then
    run_until) against the kernel source on the CPU driven through the write tables of nb_iot_environment_b200/controller.py"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'fuzz_ref_controller.py'), str(k), '120', '1tg' if k % 2 else '1'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the live reference is only present in the build container')
@pytest.mark.parametrize('k', [1, 9, 10, 11, 13])
def test_batched_wrappers_against_the_live_reference(k, oracle_lib):
    """the reference's gym wrappers over its own environment against nb_iot_environment_b200/wrappers.py over the CPU test double:
    NPRACH (discrete, default-action replay), scheduler + DiscreteActions, basic, RAR (tools/fuzz_ref_wrappers.py; kind = k % 7)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'fuzz_ref_wrappers.py'), str(k)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize('k', range(6, 16))
def test_kernel_source_on_the_cpu_against_the_oracle(k, oracle_lib):
    _drive(_Oracles, _Emu(1), k, [7100 + k, 7200 + k], 500, 0.0, 'oracle vs core (1 lane)')
    if k % 4 == 2:
        _drive(_Oracles, _Emu(32), k, [7300 + k], 150, 0.0, 'oracle vs core (32 lanes)')


@pytest.mark.gpu
@pytest.mark.parametrize('kernel', ['warp', 'tpcv'])
def test_cuda_against_the_oracle(kernel, monkeypatch):
    import __graft_entry__ as g
    g.build_cuda()
    monkeypatch.setenv('NBIOT_KERNEL', kernel)
    for k in range(10, 26):
        _, b = _drive(_Oracles, _Cuda(), k, [7400 + k, 7500 + k, 7600 + k], 400, 0.0, f'oracle vs CUDA ({kernel})')
        assert b.env.stats()['faulted'] == 0
        b.env.close()


# A cell that never reaches the external agent's state again: UE selection (Scheduling) is external, the internal RAR agent only ever
# serves CE level 2 and the NPRACH thresholds put nobody there, so nobody connects. The reference's Controller.to_next_step
# (controller.py:262-288) loops for ever on it (found by tools/fuzz_ref_controller.py, case 51); the cell must stop with a loud fault.
STALL_CONF = {'M': 2111, 'ratio': 1.0, 'buffer_range': [97, 407], 'reward_criteria': 'accumulated_delay', 'sort_criterium': 'loss', 'sc_adjustment': False, 'max_d': 1.5}
STALL_FIXED = [{'id': 2}, {'Imcs': 1, 'Nrep': 3}, {'carrier': 0, 'delay': 1, 'sc': 3}, {'carrier': 0, 'ce_level': 2, 'rar_Imcs': 1, 'delay': 1, 'sc': 3, 'Nrep': 2}, {'backoff': 0},
               {'rar_window': 4, 'mac_timer': 4, 'transmax': 1, 'panchor': 6, 'period_C0': 1, 'rep_C0': 2, 'sc_C0': 2, 'period_C1': 5, 'rep_C1': 1, 'sc_C1': 0,
                'period_C2': 1, 'rep_C2': 1, 'sc_C2': 1, 'th_C1': 2, 'th_C0': 2}]


def _stall_steps(step, records):
    from nb_iot_environment_b200.record import FAULT_STALL
    prng = np.random.default_rng(3)
    for s in range(400):
        step(np.asarray([[0, int(prng.integers(0, 4)), 3]] * 2, dtype=np.uint8))
        rec = records()
        if rec['fault'].any():
            break
    assert (rec['fault'] == FAULT_STALL).all() and s < 399
    assert (rec['state'] != 0).all()                      # stopped somewhere else than at a decision of the external agent
    t = rec['time'].copy()
    step(np.asarray([[0, 0, 3]] * 2, dtype=np.uint8))      # a faulted cell stands still
    assert np.array_equal(records()['time'], t)


@pytest.mark.parametrize('lanes', [1, '1tg'])
def test_stalled_external_step_faults_on_the_cpu_build(lanes, oracle_lib):
    from hostemu_py import EmuBatch
    from nb_iot_environment_b200.controller import ControllerTables
    tabs = ControllerTables(None, 2)
    for a, f in zip(tabs.agents, STALL_FIXED):
        a.set_action(f)
    b = EmuBatch(STALL_CONF, [9051, 9051], lanes=lanes, ctrl=tabs.build(), ue_cap=STALL_CONF['M'], **CAPS)
    b.reset()
    _stall_steps(lambda a: b.advance(a), b.records)


@pytest.mark.gpu
@pytest.mark.parametrize('kernel', ['warp', 'tpcv'])
def test_stalled_external_step_faults_instead_of_hanging(kernel, monkeypatch):
    import __graft_entry__ as g
    g.build_cuda()
    monkeypatch.setenv('NBIOT_KERNEL', kernel)
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    env = BatchedNBIoTEnv(STALL_CONF, agents_conf=None, ext_agent=2, num_envs=2, seeds=[9051, 9051], ue_cap=STALL_CONF['M'], **CAPS)
    env.reset()
    for aid, f in enumerate(STALL_FIXED):
        env.set_internal_action(aid, f)                    # DummyAgent.set_action after the Controller exists: an agent's new action reaches the shared
    env.run_until(6500)                                    # action vector when the agent next acts (controller.py:145-154) -- the NPRACH agent at t = 3000
    assert env.stats()['faulted'] == 0
    _stall_steps(lambda a: env.step(a), env.records)
    assert env.stats()['faulted'] == 2
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize('kernel', ['warp', 'tpcv'])
def test_controller_split_on_cuda_against_the_oracle(kernel, monkeypatch):
    """Controller level on the GPU: random agent set / external agent / fixed actions (set after construction, as DummyAgent.set_action is),
    observations from the device's gather kernel, rewards and records against the test-side restatement of the reference Controller"""
    import __graft_entry__ as g
    g.build_cuda()
    monkeypatch.setenv('NBIOT_KERNEL', kernel)
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.record import FAULT_STALL
    from oracle_controller import OracleController, Stalled
    from fuzz_confs import controller_case
    n_stalled = 0
    for k in range(40):
        conf, nc, cap, agents, ext, fixed, draw = controller_case(k)
        seeds = [9000 + k, 9500 + k]
        env = BatchedNBIoTEnv(conf, agents_conf=agents, ext_agent=ext, num_envs=2, seeds=seeds, n_carriers=nc, capture_effect=cap, ue_cap=conf['M'], **CAPS)
        ctrls = [OracleController(conf, s, agents_conf=agents, ext_agent=ext, n_carriers=nc, capture_effect=cap) for s in seeds]
        for aid, f in enumerate(fixed):
            env.set_internal_action(aid, f)
            for c in ctrls:
                c.t.agents[aid].set_action(f)
        obs, _ = env.reset()
        obs = obs.cpu().numpy()
        for i, c in enumerate(ctrls):
            o = c.reset()
            assert np.array_equal(o, obs[i][:len(o)]), (k, 'reset')
        names = agents[ext]['action_items']
        stalled = False
        for s in range(100):
            a = np.asarray([[draw(n) for n in names]] * 2, dtype=np.uint8)
            obs, rew, _, _, _ = env.step(a)
            obs, rew, recs = obs.cpu().numpy(), rew.cpu().numpy(), env.records()
            for i, c in enumerate(ctrls):
                try:
                    o_obs, o_rew = c.step(a[i])
                except Stalled:
                    assert recs[i]['fault'] == FAULT_STALL, (k, s, i)
                    stalled = True
                    continue
                assert np.array_equal(o_obs, obs[i][:len(o_obs)]) and o_rew == rew[i], (k, s, i)
                assert not record_diff(c.rec, recs[i], rtol=0.0), (k, s, i)
            if stalled:
                break
        n_stalled += stalled
        if not stalled:
            T = int(recs['time'].max()) + 3000
            env.run_until(T)
            recs = env.records()
            for i, c in enumerate(ctrls):
                c.run_until(T)
                assert not record_diff(c.rec, recs[i], rtol=0.0), (k, 'run_until', i)
            assert env.stats()['faulted'] == 0
        env.close()
    assert n_stalled <= 4
