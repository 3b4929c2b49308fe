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
