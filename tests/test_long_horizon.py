"""BASELINE.json configs[0]: the shipped env (single cell, default action vector) to 10^5 subframes. Compact golden of the
unmodified Python reference (tests/golden/long_cfg1.npz, tools/gen_golden_long.py): one row per decision + final counters.
The oracle is checked on CPU; the CUDA path is checked decision by decision against the golden and, at the horizon,
against the oracle record (-m gpu)."""
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, record_diff, REF_RTOL


def _load():
    z = np.load(os.path.join(GOLDEN_DIR, 'long_cfg1.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    return meta, [z[f'rows_{i}'] for i in range(len(meta['seeds']))]


def _row(rec):
    return (int(rec['time']), int(rec['state']), int(rec['total_ues']), int(rec['reward']),
            (int(rec['draws_hi']) << 32) | (int(rec['draws_lo']) & 0xFFFFFFFF))


def test_oracle_to_1e5_subframes(oracle_lib):
    from oracle.oracle_py import OracleCell
    meta, rows = _load()
    for si, seed in enumerate(meta['seeds']):
        cell = OracleCell(meta['conf'], seed)
        rec = cell.reset()
        assert _row(rec) == tuple(rows[si][0])
        for k in range(1, len(rows[si])):
            rec = cell.step(meta['action'])
            assert _row(rec) == tuple(rows[si][k]), (seed, k)
        assert rec['time'] >= meta['T']
        c = cell.counters()
        for key, v in meta['counters'][si].items():
            if key == 'av_delay':
                assert c[key] == pytest.approx(v, rel=REF_RTOL)
            else:
                assert c[key] == v, key


@pytest.mark.gpu
def test_cuda_to_1e5_subframes():
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle.oracle_py import OracleCell
    meta, rows = _load()
    n = min(len(r) for r in rows)
    env = BatchedNBIoTEnv(meta['conf'], num_envs=len(meta['seeds']), seeds=meta['seeds'], all_states_external=True, ue_cap=512, grid_w=2048)
    env.reset()
    a = np.tile(np.asarray(meta['action'], dtype=np.uint8), (len(meta['seeds']), 1))
    recs = env.records()
    for k in range(n):
        if k:
            env.node_step(a)
            if k % 16 and k < n - 1:
                continue                      # copy the records back every 16th decision (and at the end)
            recs = env.records()
        for si in range(len(meta['seeds'])):
            assert _row(recs[si]) == tuple(rows[si][k]), (si, k)
    cells = [OracleCell(meta['conf'], s) for s in meta['seeds']]
    for si, cell in enumerate(cells):
        rec = cell.reset()
        for k in range(1, n):
            rec = cell.step(meta['action'])
        assert not record_diff(rec, recs[si], rtol=0.0)
    env.close()
