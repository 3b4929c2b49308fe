"""Model-based NPRACH agent: the agent oracle against traces of the reference's NPRACH_THAgent
(CPU), the CUDA agent kernel against the oracle and the traces, and the closed loop on the device (GPU)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden', 'mb_agent.npz')


def _golden():
    z = np.load(GOLD)
    meta = json.loads(bytes(z['meta']).decode())
    return z, meta


def _info(z, p, s):
    nocc = z[f'{p}_nocc'][s]
    return {'NPRACH_detection': [z[f'{p}_det'][s, ce, :nocc[ce]].tolist() for ce in range(3)],
            'NPRACH_collision': [z[f'{p}_col'][s, ce, :nocc[ce]].tolist() for ce in range(3)],
            'distribution': z[f'{p}_dist'][s].tolist(), 'incoming': [0.0] * int(z[f'{p}_ninc'][s])}


def _prefixes(meta):
    return [(f'loop_{i}', m['no_th']) for i, m in enumerate(meta['loops'])] + [(f'synth_{i}', m['no_th']) for i, m in enumerate(meta['synth'])]


def test_agent_oracle_matches_reference_traces():
    from oracle.mb_agent_oracle import OracleNPRACHAgent
    z, meta = _golden()
    branches = set()
    for p, no_th in _prefixes(meta):
        ag = OracleNPRACHAgent(no_th=no_th)
        for s in range(len(z[f'{p}_out'])):
            ag.set_parameter(float(z[f'{p}_margin'][s]))
            k_before = ag.k
            out = ag.get_action(_info(z, p, s))
            assert out == z[f'{p}_out'][s].tolist(), (p, s)
            assert ag.lambda_estimate == z[f'{p}_lam'][s] and ag.k == z[f'{p}_k'][s], (p, s)      # floats bit for bit
            assert ag.msg3_detection == z[f'{p}_m3'][s].tolist(), (p, s)
            branches.add('gather' if k_before < 100 else 'no_th' if no_th else 'table')
            if ag.conf != [2, 2, 2, 2, 2, 2]:
                branches.add('memory')
    assert branches == {'gather', 'no_th', 'table', 'memory'}       # every branch of get_action is in the traces


def test_agent_tables_are_consistent():
    from oracle.mb_agent_oracle import tables
    t = tables()
    assert t['ml_est'].shape == (4, 128, 64) and t['rate_conf'].shape == (186, 8)
    # rows N = 3, 6, 9 never consult the model (no probabilities for r < 12): max(0, d - 1)
    d = np.arange(128)[:, None]
    for ni in range(3):
        assert (t['ml_est'][ni] == np.maximum(0, d - 1)).all()
    assert (t['ml_est'][3, :13] != np.maximum(0, d[:13] - 1)).any()
    for ce in range(3):
        n = int(t['cfg_n'][ce])
        assert (np.diff(t['cfg_rate'][ce, :n]) >= 0).all() and np.isinf(t['cfg_rate'][ce, n:]).all()


# ------------------------------------------------------------------ GPU: the CUDA agent through the C ABI
def _write_records(env, rows):
    """overwrite the decision records of the batch with the given structured rows (the state buffer is caller-owned)"""
    import torch
    from nb_iot_environment_b200.record import RECORD_DTYPE
    raw = torch.from_numpy(np.frombuffer(rows.tobytes(), dtype=np.uint8).copy())
    env.state[env._rec_off:env._rec_off + rows.size * RECORD_DTYPE.itemsize] = raw.to(env.state.device)


@pytest.mark.gpu
@pytest.mark.parametrize('no_th', [False, True])
def test_agent_kernel_matches_reference_traces_and_oracle(no_th):
    """every get_action of the golden traces, replayed through nbiot_mb_agent_act: actions, arrival estimate, sample
    counter and msg3 detection estimate bit-identical to the reference agent"""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.mb_agent import BatchedNPRACHAgent
    from nb_iot_environment_b200.record import RECORD_DTYPE
    z, meta = _golden()
    pref = [p for p, f in _prefixes(meta) if f == no_th]
    N = len(pref)
    env = BatchedNBIoTEnv({'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}, num_envs=N, seeds=list(range(N)),
                          ue_cap=64, grid_w=256, hist_cap=64)
    env.reset()
    agent = BatchedNPRACHAgent(env, no_th=no_th)
    for s in range(40):
        rows = np.zeros(N, dtype=RECORD_DTYPE)
        for i, p in enumerate(pref):
            rows[i]['state'] = 1
            rows[i]['det_hist'] = z[f'{p}_det'][s]; rows[i]['col_hist'] = z[f'{p}_col'][s]; rows[i]['n_occ'] = z[f'{p}_nocc'][s]
            rows[i]['distribution'] = z[f'{p}_dist'][s]; rows[i]['n_incoming'] = z[f'{p}_ninc'][s]
        _write_records(env, rows)
        agent.set_parameter(np.asarray([z[f'{p}_margin'][s] for p in pref]))
        act = agent.get_action().cpu().numpy()
        lam, k, m3 = agent.lambda_estimate.cpu().numpy(), agent.k.cpu().numpy(), agent.msg3_detection.cpu().numpy()
        for i, p in enumerate(pref):
            assert act[i].tolist() == z[f'{p}_out'][s].tolist(), (p, s)
            assert lam[i] == z[f'{p}_lam'][s] and k[i] == z[f'{p}_k'][s] and m3[i].tolist() == z[f'{p}_m3'][s].tolist(), (p, s)
    # a cell that is not at an NPRACH_update decision is left alone
    rows['state'] = 0
    _write_records(env, rows)
    before = agent.state.clone()
    agent.get_action()
    assert torch.equal(before, agent.state) and agent.faults() == 0
    agent.close(); env.close()


@pytest.mark.gpu
def test_agent_closed_loop_on_device_matches_reference():
    """evaluate_MBRL.py set-up entirely on the device (environment -> records -> agent -> actions -> environment) against the
    reference environment + agent + NPRACH_agent_wrapper: same configurations, estimates, observations, rewards, samples"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_gpu_parity import NPRACH_AGENTS
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.mb_agent import BatchedNPRACHAgent
    from nb_iot_environment_b200.wrappers import NPRACHAgentWrapperTraces as NPRACHAgentWrapper       # the plain wrapper + running means
    z, meta = _golden()
    for ci, m in enumerate(meta['loops']):
        p = f'loop_{ci}'
        env = BatchedNBIoTEnv(m['conf'], agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=2, seeds=[m['seed'], m['seed']], ue_cap=1000, grid_w=2048, hist_cap=1024)
        agent = BatchedNPRACHAgent(env, no_th=m['no_th'])
        w = NPRACHAgentWrapper(env, agent, n_actions=26, obs_items=[0, 1, 2, 3, 4, 5, 9, 10, 11, 12], obs_window=1, bounds=[0.3, 2.0], discrete=True, norm=100)
        obs, _ = w.reset()
        np.testing.assert_allclose(obs.cpu().numpy()[0], z[f'{p}_obs0'], rtol=1e-9, atol=1e-12)
        for s in range(m['steps']):
            a = int(z[f'{p}_act'][s])
            obs, r, _, _, _ = w.step(np.asarray([a, a]))
            conf = w.conf.cpu().numpy()
            assert conf[0].tolist() == z[f'{p}_out'][s].tolist() and conf[1].tolist() == conf[0].tolist(), (p, s)
            assert agent.lambda_estimate.cpu().numpy()[0] == z[f'{p}_lam'][s], (p, s)
            assert r.cpu().numpy()[0] == z[f'{p}_rew'][s], (p, s)
            np.testing.assert_allclose(obs.cpu().numpy()[0], z[f'{p}_obs'][s], rtol=1e-9, atol=1e-12, err_msg=f'{p} step {s}')
        for name, key in (('departures', 's_departures'), ('NPRACH_occupation', 's_occupation'), ('service_times', 's_service'), ('beta', 's_beta')):
            got = np.asarray([x.cpu().numpy()[0] for x in agent.samples[name]])
            np.testing.assert_allclose(got, z[f'{p}_{key}'], rtol=1e-9, atol=1e-12, err_msg=f'{p} samples {name}')
            avg = 0.0                                            # NPRACH_agent_wrapper_traces.step (wrappers.py:337-340) on the reference's samples
            for n, sample in enumerate(z[f'{p}_{key}'], start=1):
                avg += (float(sample) - avg) / n
            np.testing.assert_allclose(w.averages[name].cpu().numpy()[0], avg, rtol=1e-9, atol=1e-12, err_msg=f'{p} running mean {name}')
        assert agent.faults() == 0
        agent.close(); env.close()


@pytest.mark.skipif(not os.path.isdir(os.environ.get('NBIOT_REFERENCE', '/root/reference')), reason='the live reference is only present in the build container')
def test_agent_oracle_against_the_live_reference_agent():
    """random inputs (tests/fuzz_agent.py) through the reference's NPRACH_THAgent and through its restatement, in a process of its own"""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'fuzz_ref_agent.py'), '0', '12'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_agent_kernel_on_random_inputs_against_the_oracle():
    """48 cells, each fed its own seeded random sequence of NPRACH_update records (idle to far overloaded, both agent variants):
    the CUDA agent's actions, arrival estimate, sample counter and msg3 estimate against the restatement, floats bit for bit"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from fuzz_agent import agent_sequence
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.mb_agent import BatchedNPRACHAgent
    from nb_iot_environment_b200.record import RECORD_DTYPE
    from oracle.mb_agent_oracle import OracleNPRACHAgent
    for no_th in (False, True):
        seqs = [s for f, s in (agent_sequence(k) for k in range(100, 200)) if f == no_th][:48]
        N = len(seqs)
        env = BatchedNBIoTEnv({'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}, num_envs=N, seeds=list(range(N)),
                              ue_cap=64, grid_w=256, hist_cap=64)
        env.reset()
        agent = BatchedNPRACHAgent(env, no_th=no_th)
        orcs = [OracleNPRACHAgent(no_th=no_th) for _ in range(N)]
        for s in range(60):
            rows = np.zeros(N, dtype=RECORD_DTYPE)
            for i, q in enumerate(seqs):
                info = q[s]
                rows[i]['state'] = 1
                for ce in range(3):
                    d, c = info['NPRACH_detection'][ce], info['NPRACH_collision'][ce]
                    rows[i]['n_occ'][ce] = len(d); rows[i]['det_hist'][ce][:len(d)] = d; rows[i]['col_hist'][ce][:len(c)] = c
                rows[i]['distribution'] = info['distribution']; rows[i]['n_incoming'] = len(info['incoming'])
            _write_records(env, rows)
            agent.set_parameter(np.asarray([q[s]['margin'] for q in seqs]))
            act = agent.get_action().cpu().numpy()
            lam, k, m3 = agent.lambda_estimate.cpu().numpy(), agent.k.cpu().numpy(), agent.msg3_detection.cpu().numpy()
            for i, q in enumerate(seqs):
                orcs[i].set_parameter(q[s]['margin'])
                want = orcs[i].get_action(q[s])
                assert act[i].tolist() == want, (no_th, i, s)
                assert lam[i] == orcs[i].lambda_estimate and k[i] == orcs[i].k and m3[i].tolist() == [float(x) for x in orcs[i].msg3_detection], (no_th, i, s)
        assert agent.faults() == 0
        agent.close(); env.close()
