"""The batched wrappers (nb_iot_environment_b200/wrappers.py) against traces of the REFERENCE's wrappers running on the
reference's own Controller + gym Environment (tests/golden/wrappers_*.npz, tools/gen_golden_wrappers.py):
on a CPU test double built from the oracle (always) and on the CUDA environment (-m gpu)."""
import json

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, REF_RTOL
from test_gpu_parity import NPRACH_AGENTS, NPUSCH_AGENTS

import os


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    S = len(meta['seeds'])
    return meta, [z[f'acts_{i}'] for i in range(S)], [z[f'obs_{i}'] for i in range(S)], [z[f'rew_{i}'] for i in range(S)]


def _check_nprach(make_env, traces=False):
    from nb_iot_environment_b200.wrappers import NPRACHWrapper, NPRACHWrapperTraces
    meta, acts, obs, rew = _load('wrappers_nprach')
    base = make_env(meta['conf'], NPRACH_AGENTS, 3, meta['seeds'])
    env = NPRACHWrapperTraces(base, meta['metrics']) if traces else NPRACHWrapper(base)
    assert env.action_nvec.tolist() == [105, 4, 4, 4, 6, 6, 6]
    o, _ = env.reset()
    o = o.cpu().numpy()
    for i in range(len(meta['seeds'])):
        assert np.allclose(o[i], obs[i][0], rtol=REF_RTOL, atol=1e-12)
    n_invalid = 0
    for k in range(acts[0].shape[0]):
        a = np.stack([acts[i][k] for i in range(len(acts))])
        o, r, term, trunc, _ = env.step(a)
        o, r = o.cpu().numpy(), r.cpu().numpy()
        n_invalid += int((~env.last_valid).sum())
        for i in range(len(acts)):
            assert np.allclose(o[i], obs[i][k + 1], rtol=REF_RTOL, atol=1e-12), (k, i)
            assert r[i] == pytest.approx(rew[i][k], rel=1e-12), (k, i)
    assert n_invalid > 0          # the trace exercises the replay-previous-action branch
    if traces:                    # NPRACH_wrapper_traces.samples (wrappers.py:210-255): info[metric], lists as their mean
        z = np.load(os.path.join(GOLDEN_DIR, 'wrappers_nprach.npz'))
        for metric in meta['metrics']:
            got = env.samples_array(metric).cpu().numpy()
            for i in range(len(meta['seeds'])):
                assert np.allclose(got[:, i], z[f'samples_{metric}_{i}'], rtol=REF_RTOL, atol=1e-12), (metric, i)


def _check_npusch(make_env):
    from nb_iot_environment_b200.wrappers import BasicSchedulerWrapper, DiscreteActions
    meta, acts, obs, rew = _load('wrappers_npusch')
    env = DiscreteActions(BasicSchedulerWrapper(make_env(meta['conf'], NPUSCH_AGENTS, 0, meta['seeds'])))
    assert env.n_actions == 4 * 7 * 5
    o, _ = env.reset()
    n_impossible = 0
    for k in range(acts[0].shape[0]):
        a = np.stack([acts[i][k] for i in range(len(acts))])
        o, r, _, _, _ = env.step(a)
        o, r = o.cpu().numpy(), r.cpu().numpy()
        n_impossible += int(env.env.last_impossible.sum())
        for i in range(len(acts)):
            assert np.allclose(o[i], obs[i][k + 1], rtol=REF_RTOL, atol=1e-12), (k, i)
            assert r[i] == pytest.approx(rew[i][k], rel=1e-12), (k, i)
    assert n_impossible > 0       # ... and the answer-without-stepping branch


def _check_basic(make_env):
    """BasicWrapper of the reference (wrappers.py:47-82: UE selection only, scalar action) over agent 0 of the six default agents"""
    from nb_iot_environment_b200.wrappers import BasicWrapper
    meta, acts, obs, rew = _load('wrappers_basic')
    env = BasicWrapper(make_env(meta['conf'], None, 0, meta['seeds']))
    assert list(env.action_nvec) == [4]
    env.reset()
    n_impossible = 0
    for k in range(acts[0].shape[0]):
        a = np.asarray([[acts[i][k]] for i in range(len(acts))])
        o, r, _, _, _ = env.step(a)
        o, r = o.cpu().numpy(), r.cpu().numpy()
        n_impossible += int(env.last_impossible.sum())
        for i in range(len(acts)):
            assert np.allclose(o[i], obs[i][k + 1], rtol=REF_RTOL, atol=1e-12), (k, i)
            assert r[i] == pytest.approx(rew[i][k], rel=1e-12), (k, i)
    assert n_impossible > 0


def _check_rar(make_env):
    """RAR_wrapper of the reference (wrappers.py:366-421) on its own environment, both reward criteria, against RARWrapper"""
    from nb_iot_environment_b200.wrappers import RARWrapper
    z = np.load(os.path.join(GOLDEN_DIR, 'wrappers_rar.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    seen = set()
    for si, (seed, crit) in enumerate(zip(meta['seeds'], meta['criteria'])):
        env = RARWrapper(make_env(meta['conf'], NPRACH_AGENTS, 1, [seed]), reward_criteria=crit)
        env.reset()
        acts, rew = z[f'acts_{si}'], z[f'rew_{si}']
        for k in range(acts.shape[0]):
            _, r, _, _, _ = env.step(acts[k][None, :])
            assert float(r[0]) == rew[k], (crit, k)
            seen.add(float(rew[k]))
    assert -1.0 in seen and 1.0 in seen and len(seen) >= 5      # penalties, saturated and fractional rewards all occur


def _oracle_env(conf, agents, ext, seeds):
    from oracle_vector_env import OracleVectorEnv
    return OracleVectorEnv(conf, agents, ext, seeds)


def _cuda_env(conf, agents, ext, seeds):
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    return BatchedNBIoTEnv(conf, agents_conf=agents, ext_agent=ext, num_envs=len(seeds), seeds=seeds, ue_cap=conf['M'], grid_w=2048, hist_cap=1024)


def test_validity_table_and_discrete_decode():
    from nb_iot_environment_b200 import wrappers as W, parameters as par
    t = W.nprach_validity_table()
    assert t.shape == (105, 4, 4, 4, 6, 6, 6) and 0 < t.mean() < 1
    # parameters.compute_NPRACH_sf for the default thresholds (-106, -100): no repetitions for CE0/CE1, 8 for CE2
    assert W.compute_NPRACH_sf(-106, -100) == [6, 6, 45]
    assert W.compute_NPRACH_sf(-115, 0) == [0, 45, 45]
    d = W.to_discrete([4, 7, 5])
    assert d.shape == (140, 3) and d[0].tolist() == [0, 0, 0] and d[1].tolist() == [0, 0, 1] and d[-1].tolist() == [3, 6, 4]
    assert len(par.th_pairs) == 105


def test_nprach_wrapper_on_oracle(oracle_lib):
    _check_nprach(_oracle_env)


def test_scheduler_wrapper_on_oracle(oracle_lib):
    _check_npusch(_oracle_env)


def test_basic_wrapper_on_oracle(oracle_lib):
    _check_basic(_oracle_env)


@pytest.mark.gpu
def test_basic_wrapper_on_cuda():
    _check_basic(_cuda_env)


def test_rar_wrapper_on_oracle(oracle_lib):
    _check_rar(_oracle_env)


@pytest.mark.gpu
def test_rar_wrapper_on_cuda():
    _check_rar(_cuda_env)


@pytest.mark.gpu
def test_nprach_wrapper_on_cuda():
    _check_nprach(_cuda_env)


@pytest.mark.gpu
def test_nprach_wrapper_traces_on_cuda():
    _check_nprach(_cuda_env, traces=True)


@pytest.mark.gpu
def test_scheduler_wrapper_on_cuda():
    _check_npusch(_cuda_env)


@pytest.mark.gpu
def test_rar_wrapper_cuda_matches_oracle():
    from nb_iot_environment_b200.wrappers import RARWrapper
    conf = {'ratio': 0.5, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'ce_mcs_automatic': False}
    agents = NPRACH_AGENTS
    seeds = [600, 601, 602]
    a_env, b_env = RARWrapper(_cuda_env(conf, agents, 1, seeds)), RARWrapper(_oracle_env(conf, agents, 1, seeds))
    a_env.reset(); b_env.reset()
    rng = np.random.default_rng(2)
    for _ in range(300):
        a = np.stack([rng.integers(0, 3, 3), rng.integers(0, 3, 3), rng.integers(0, 5, 3)], axis=1)
        oa, ra, _, _, _ = a_env.step(a)
        ob, rb, _, _, _ = b_env.step(a)
        assert np.array_equal(oa.cpu().numpy(), ob.numpy()) and np.array_equal(ra.cpu().numpy(), rb.numpy())
