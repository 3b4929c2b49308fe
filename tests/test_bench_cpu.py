"""Host logic of bench.py that needs no GPU: the action stream is a function of (seed, step, global cell index), and the CPU arm really
is the unmodified reference -- its observations and rewards, produced one environment per process, equal the oracle's for the same cells."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_actions_do_not_depend_on_steps_ranks_or_batch_size():
    import bench
    a = bench.ext_actions(6, 32, seed=1234)
    assert np.array_equal(bench.ext_actions(3, 32, seed=1234), a[:3])                         # --steps
    assert np.array_equal(bench.ext_actions(6, 8, seed=1234, first_cell=16), a[:, 16:24])     # ranks / shards
    assert not np.array_equal(bench.ext_actions(6, 32, seed=1235), a)
    from nb_iot_environment_b200 import parameters as par
    pairs = {tuple(par.th_indexes[p]) for p in par.th_pairs}
    assert all(tuple(x) in pairs for x in a[..., 0:2].reshape(-1, 2)) and a[..., 2:5].max() == 3 and a[..., 5:8].max() == 5
    u = bench.uniform_actions(77, 4, 10, 5, [4, 7, 5])
    assert np.array_equal(u, bench.uniform_actions(77, 4, 0, 15, [4, 7, 5])[10:]) and (u.max(axis=0) < [4, 7, 5]).all()


def test_reference_vector_env_is_the_reference(oracle_lib):
    """oracle/ref_vector_env.py (the CPU arm of bench.py): two reference environments in two processes, stepped with the benchmark's
    actions, against a test-side Controller over the C oracle for the same seeds -- rewards equal, observations to 1e-9."""
    from oracle import ref_vector_env as rv
    if rv.reference_dir() is None:
        pytest.skip('no staged reference (oracle/_ref) and no /root/reference on this box')
    import bench
    from oracle_controller import OracleController
    P, W, K = 2, 1, 3
    acts = bench.ext_actions(W + K, P, seed=1234)
    r = rv.run(bench.CONF, bench.AGENTS, bench.EXT_AGENT, list(range(P)), acts, W, P)
    assert r['cores'] == P and r['subframes'] == P * K * bench.SF_PER_STEP and r['value'] > 0
    for i in range(P):
        c = OracleController(bench.CONF, i, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT)
        c.reset()
        for s in range(W + K):
            obs, rew = c.step(acts[s, i])
            o_ref, r_ref = r['traces'][i][s]
            assert rew == r_ref and np.allclose(obs, o_ref, rtol=1e-9, atol=1e-12), (i, s)
