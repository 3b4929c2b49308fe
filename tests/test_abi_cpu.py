"""CPU-side checks of the product boundary: the CUDA library loads without a GPU, exports every symbol
include/nbiot.h declares, its host random twins equal the oracle's, and nothing in the package reaches
for the oracle or a CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cuda_lib():
    import __graft_entry__ as g
    g.build_cuda()
    from nb_iot_environment_b200 import _cabi
    return _cabi.load()


def test_every_declared_symbol_is_exported(cuda_lib):
    from nb_iot_environment_b200 import _cabi
    text = open(os.path.join(ROOT, 'include', 'nbiot.h')).read() + open(os.path.join(ROOT, 'include', 'nbiot_agent.h')).read()
    declared = sorted(set(re.findall(r'^(?:int|size_t|double|void|uint64_t|long long|const char \*)\s*\*?(nbiot_[a-z0-9_]+)\s*\(', text, flags=re.M)))
    assert declared == sorted(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(cuda_lib, name), name


def test_struct_sizes_match(cuda_lib, oracle_lib):
    from nb_iot_environment_b200.record import RECORD_DTYPE, NbiotConf
    from nb_iot_environment_b200.controller import NbiotCtrl
    from nb_iot_environment_b200._cabi import ObsItem
    assert oracle_lib.orc_record_size() == RECORD_DTYPE.itemsize == 1632
    assert oracle_lib.orc_conf_size() == ctypes.sizeof(NbiotConf)
    assert ctypes.sizeof(NbiotCtrl) == 8 + 24 + 24 + 96 + 24 + 96
    assert ctypes.sizeof(ObsItem) == 24
    from nb_iot_environment_b200._cabi import MbTables, MB_STATE_BYTES
    assert ctypes.sizeof(MbTables) == 240 and MB_STATE_BYTES == 56          # nbiot_mb_tables_t / nbiot_mb_state_t (include/nbiot_agent.h)
    assert cuda_lib.nbiot_hdr_bytes() % 16 == 0


def test_host_twins_equal_oracle_twins(cuda_lib, oracle_lib):
    xy1, xy2 = (ctypes.c_double * 2)(), (ctypes.c_double * 2)()
    for seed in (0, 7, 2 ** 40 + 3):
        for ctr in (0, 1, 12345, 2 ** 33):
            assert cuda_lib.nbiot_rng_u01(seed, ctr, 0) == oracle_lib.orc_rng_u01(seed, ctr, 0)
            assert cuda_lib.nbiot_rng_normal(seed, ctr, 0) == oracle_lib.orc_rng_normal(seed, ctr, 0)
            assert cuda_lib.nbiot_rng_bounded(seed, ctr, 1, 524288) == oracle_lib.orc_rng_bounded(seed, ctr, 1, 524288)
            cuda_lib.nbiot_rng_pair(seed, ctr, 0, xy1)
            oracle_lib.orc_rng_pair(seed, ctr, 0, xy2)
            assert tuple(xy1) == tuple(xy2)


def test_no_gpu_means_loud_failure(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
    with pytest.raises(_cabi.NbiotError):
        BatchedNBIoTEnv({'ratio': 1.0, 'M': 100, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'})
    # the C entry point itself refuses as well
    from nb_iot_environment_b200.record import make_conf
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.tables import load_tables
    conf = make_conf({'ratio': 1.0, 'M': 100, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'})
    ctrl = ControllerTables().build()
    lut, mcs = load_tables()
    seeds = np.zeros(1, dtype=np.uint64)
    h = ctypes.c_void_p()
    dummy = np.zeros(16, dtype=np.uint8)
    rc = cuda_lib.nbiot_create(ctypes.byref(conf), 1, 0, dummy.ctypes.data, 16, seeds.ctypes.data, lut.ctypes.data, mcs.ctypes.data,
                               ctypes.byref(ctrl), None, ctypes.byref(h))
    assert rc < 0 and cuda_lib.nbiot_last_error()


def test_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'nb_iot_environment_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle_py' not in text and 'libnbiot_oracle' not in text and '_hostemu' not in text.replace('tests/_hostemu', ''), f


def test_conf_validation():
    from nb_iot_environment_b200.record import make_conf
    base = {'ratio': 1.0, 'M': 100, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
    make_conf(base)
    c = make_conf(dict(base, clustered=True, cluster_ues=[3, 8], cluster_prob=0.9, cluster_range=2.0, statistics=True))
    assert (c.clustered, c.cluster_lo, c.cluster_hi, c.cluster_prob, c.cluster_range, c.stat_cap) == (1, 3, 8, 0.9, 2.0, 2048)
    assert make_conf(base).stat_cap == 0 and make_conf(base, stat_cap=64).stat_cap == 64
    with pytest.raises(ValueError):
        make_conf(dict(base, clustered=True, cluster_ues=[5, 5]))
    c = make_conf(dict(base, traces=True))                          # perf_monitor.py:61-69: the NPRACH and msg3 sample lists
    assert (c.trace_mask, c.trace_cap, c.stat_cap) == (3, 8192, 0)
    c = make_conf(dict(base, statistics=True), trace_cap=100)
    assert (c.trace_mask, c.trace_cap, c.stat_cap) == (0x7C, 100, 2048) and make_conf(base).trace_cap == 0
    with pytest.raises(NotImplementedError):
        make_conf(dict(base, animate_stats=True))
    with pytest.raises(NotImplementedError):
        make_conf(dict(base, reward_criteria='average_total_delay'))
    with pytest.raises(KeyError):
        make_conf(dict(base, bogus=1))
    with pytest.raises(ValueError):
        make_conf(base, grid_w=1000)
    for bad in (dict(buffer_range=[100, 70000]), dict(buffer_range=[0, 10]), dict(M=0), dict(ratio=-0.1)):
        with pytest.raises(ValueError):
            make_conf(dict(base, **bad))
    assert (make_conf(base).kernel, make_conf(base, kernel='warp').kernel, make_conf(base, kernel='thread_voted').kernel) == (0, 1, 3)      # NBIOT_KERNEL_*
    with pytest.raises(ValueError):
        make_conf(base, kernel='fastest')
    assert make_conf(dict(base, buffer_range=[300, 200])).buffer_lo == 300          # legal in the reference: every buffer is 300 (ue_generator.py:98-104)


def test_controller_tables_follow_reference_split():
    from nb_iot_environment_b200 import parameters as par
    from nb_iot_environment_b200.controller import ControllerTables
    t = ControllerTables()                      # the reference's six default agents
    assert t.initial_action()[:23] == [0, 1, 1, 0, 0, 3, 1, 3, 6, 2, 4, 12, 3, 0, 1, 2, 0, 1, 3, 2, 1, 8, 11]
    c = t.build()
    assert c.ext_state_mask == 0 and c.ext_k == 0
    # Scheduling: agents 0,1,2 write id, Imcs, Nrep, carrier, delay, sc
    assert [c.state_writes[0][i] for i in range(7)] == [0, 0, 2, -1, 0, 3, 1]
    # RAR_window: agent 3 writes carrier, ce_level, rar_Imcs, delay, sc, Nrep
    assert [c.state_writes[2][i] for i in range(7)] == [0, 1, 1, -1, 0, 3, 1]
    t.set_ext_agent(0)
    c = t.build()
    assert c.ext_state_mask == 1 and c.ext_k == 1 and c.ext_a_mask[0] == par.control_items['id']
    # agent 0's chain: next=1 (Imcs, Nrep), then 2 (carrier, delay, sc)
    assert [c.ext_post[i] for i in range(7)] == [0, -1, 2, -1, 0, 3, 1]
    assert all(c.state_writes[0][i] == -1 for i in range(23))      # nobody acts before agent 0 in Scheduling
    assert t.obs_dim == 1 + 4 * 4 + 40
