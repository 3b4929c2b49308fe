"""Parity of the CUDA path, called through the C ABI, with the CPU oracle and with the golden traces of
the unmodified Python reference. Integer fields bit-exact; float fields bit-exact against the oracle
(both evaluate include/nbiot_math.h) and within 1e-9 relative of the reference's numpy arithmetic
(the contract in BASELINE.json is 1e-5)."""
import numpy as np
import pytest

from golden_util import CASES, Golden, record_diff, assert_pm_lists, REF_RTOL

pytestmark = pytest.mark.gpu

CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
CAPS = {'capture_effect': dict(grid_w=2048), 'capture_effect_two_carriers': dict(grid_w=2048), 'traces_two_carriers': dict(grid_w=2048), 'manual_branches': dict(grid_w=4096), 'cfg3_npusch_random_policy': dict(grid_w=2048), 'cfg2_random_policy': dict(grid_w=2048),
        'cfg4_two_carriers_bursty': dict(grid_w=2048)}

# agent sets of the two shipped scenarios (scripts/nprach/evaluate_RL.py:46-77, scripts/npusch/experiments_RL.py:40-78)
NPRACH_AGENTS = [
    {'id': 0, 'action_items': ['id', 'Imcs', 'Nrep', 'carrier', 'delay', 'sc'], 'obs_items': [], 'next': -1, 'states': ['Scheduling']},
    {'id': 1, 'action_items': ['ce_level', 'rar_Imcs', 'Nrep'], 'obs_items': [], 'next': -1, 'states': ['RAR_window']},
    {'id': 2, 'action_items': ['rar_window', 'mac_timer', 'transmax', 'panchor', 'backoff'], 'obs_items': [], 'next': -1, 'states': ['RAR_window_end']},
    {'id': 3, 'action_items': ['th_C1', 'th_C0', 'sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2'],
     'obs_items': ['detection_ratios', 'colision_ratios', 'msg3_detection', 'NPRACH_occupation', 'av_delay', 'distribution'],
     'next': -1, 'states': ['NPRACH_update']},
]
NPUSCH_AGENTS = [
    {'id': 0, 'action_items': ['id', 'Imcs', 'Nrep'], 'obs_items': ['total_ues', 'connection_time', 'loss', 'sinr', 'buffer', 'carrier_state'],
     'next': 1, 'states': ['Scheduling']},
    {'id': 1, 'action_items': ['carrier', 'delay', 'sc'], 'obs_items': [], 'next': -1, 'states': ['Scheduling']},
    {'id': 2, 'action_items': ['carrier', 'ce_level', 'rar_Imcs', 'delay', 'sc', 'Nrep'], 'obs_items': [], 'next': -1, 'states': ['RAR_window']},
    {'id': 3, 'action_items': ['backoff'], 'obs_items': [], 'next': -1, 'states': ['RAR_window_end']},
    {'id': 4, 'action_items': ['th_C1', 'th_C0', 'sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2'], 'obs_items': [],
     'next': -1, 'states': ['NPRACH_update']},
]
CFG3 = {'M': 2000, 'ratio': 1, 'buffer_range': [100, 500], 'reward_criteria': 'average_delay', 'sc_adjustment': True, 'mcs_automatic': False}


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


@pytest.fixture(params=['warp', 'tpc', 'tpcv'])
def kernel(request, monkeypatch):
    """the simulation kernels: one warp per cell (csrc/nbiot_simk.cuh), one thread per cell and one thread per cell with voted
    dispatch (csrc/nbiot_simt.cuh); without the knob nbiot_create picks by batch size"""
    monkeypatch.setenv('NBIOT_KERNEL', request.param)
    return request.param


@pytest.mark.parametrize('case', CASES)
def test_golden_traces_and_oracle(case, kernel):
    """NodeB.step-level replay of the reference's traces: every decision record of every seed."""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle.oracle_py import OracleCell
    g = Golden(case)
    caps = dict(ue_cap=g.conf['M'], grid_w=1024, hist_cap=1024)
    caps.update(CAPS.get(case, {}))
    env = BatchedNBIoTEnv(g.conf, num_envs=len(g.seeds), seeds=g.seeds, n_carriers=g.n_carriers, all_states_external=True, capture_effect=g.capture_effect, **caps)
    cells = [OracleCell(g.conf, s, n_carriers=g.n_carriers, capture_effect=g.capture_effect) for s in g.seeds]
    env.reset()
    recs = env.records()
    hists = [g.hist(si) for si in range(len(g.seeds))]
    hidx = [0] * len(g.seeds)
    for d in range(-1, g.decisions):
        if d >= 0:
            env.node_step(g.actions[:, d, :])
            recs = env.records()
        lists = None
        for si, cell in enumerate(cells):
            orec = cell.reset() if d < 0 else cell.step(g.actions[si, d])
            diff = record_diff(orec, recs[si], rtol=0.0)
            assert not diff, f'{case} seed {g.seeds[si]} decision {d + 1} vs oracle: {diff[:3]}'
            diff = record_diff(g.records[si, d + 1], recs[si], rtol=REF_RTOL)
            assert not diff, f'{case} seed {g.seeds[si]} decision {d + 1} vs reference: {diff[:3]}'
            if recs[si]['state'] == 1:
                if lists is None:
                    lists = env.period_lists()
                ginc, gdl, gsv = hists[si][hidx[si]]
                hidx[si] += 1
                assert np.array_equal(lists[1][si, :len(gdl)], gdl) and np.array_equal(lists[2][si, :len(gsv)], gsv)
                assert np.allclose(lists[0][si, :len(ginc)], ginc, rtol=REF_RTOL, atol=0)
    assert not (recs['fault'] & 0x1F).any()
    if g.pm_lists(0) is not None:
        # conf['traces'] / conf['statistics']: PerfMonitor's sample lists (perf_monitor.py:61-79) through env.traces() / env.histories()
        tr = env.traces()
        hs = env.histories() if g.conf.get('statistics') else [{} for _ in g.seeds]
        for si in range(len(g.seeds)):
            got = dict(tr[si]); got.update(hs[si])
            assert not tr[si]['truncated'] and not hs[si].get('samples_truncated', False)
            assert_pm_lists(got, g.pm_lists(si), f'{case} seed {g.seeds[si]}')
    if g.stats(0) is not None:
        # conf['statistics']: the per-departure histories (perf_monitor.py:61-84) and the batch histogram computed on the device
        hs = env.histories()
        all_delays = []
        for si in range(len(g.seeds)):
            gs = g.stats(si)
            assert hs[si]['served_users'] == len(gs['delay']) and not hs[si]['truncated']
            for k in ('delay', 'connection', 'access'):
                assert np.array_equal(hs[si][k], gs[k]), k
            assert np.array_equal(hs[si]['th'], gs['th'])
            all_delays.append(gs['delay'])
        d = np.concatenate(all_delays)
        want = np.bincount(np.minimum(d // 250, 31), minlength=32)
        assert np.array_equal(env.stat_histogram('delay', bin_width=250, n_bins=32).cpu().numpy(), want)
    env.close()


@pytest.mark.gpu
def test_statistics_ring_keeps_the_latest_departures():
    """more departures than stat_cap: the ring holds the last stat_cap of them, in order, and says so"""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    conf = dict(CFG1, statistics=True)
    full = BatchedNBIoTEnv(conf, num_envs=3, seeds=[5, 6, 7], ue_cap=512, grid_w=1024)
    small = BatchedNBIoTEnv(conf, num_envs=3, seeds=[5, 6, 7], ue_cap=512, grid_w=1024, stat_cap=16)
    for e in (full, small):
        e.reset(); e.run_until(20000)
    hf, hs = full.histories(), small.histories()
    for i in range(3):
        n = hf[i]['served_users']
        assert n > 16 and hs[i]['served_users'] == n and hs[i]['truncated'] and not hf[i]['truncated']
        for k in ('delay', 'connection', 'access', 'th'):
            assert np.array_equal(hs[i][k], hf[i][k][-16:]), k
    small.reset()
    assert small.stat_counts().cpu().numpy().tolist() == [0, 0, 0]          # perf_monitor.reset() restarts the histories
    no = BatchedNBIoTEnv(CFG1, num_envs=1, ue_cap=64, grid_w=256)
    with pytest.raises(Exception):
        no.histories()
    for e in (full, small, no):
        e.close()


@pytest.mark.gpu
def test_grid_in_place_equals_staged_grid(monkeypatch):
    """short launches (an external scheduling agent) leave the look-ahead ring in HBM, long ones stage it in shared memory: both
    launch paths, and both register budgets of the kernel, give bit-identical states"""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    rng = np.random.default_rng(3)
    states = []
    for stage, build in (('1', 'lo'), ('0', 'lo'), ('0', 'hi'), (None, None)):
        for k, v in (('NBIOT_STAGE_GRID', stage), ('NBIOT_SIM_BUILD', build)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=96, seeds=list(range(96)), ue_cap=512, grid_w=2048, hist_cap=64)
        env.reset()
        r = np.random.default_rng(3)
        for _ in range(150):
            env.step(_random_ext_actions(r, env.action_nvec, 96))
        env.run_until(9000)                                     # a long launch on the internal agents on top
        torch.cuda.synchronize()
        states.append(env.state.cpu().numpy().copy())
        assert env.stats()['faulted'] == 0
        env.close()
    for s in states[1:]:
        assert np.array_equal(states[0], s)


def test_thread_per_cell_equals_warp_per_cell(monkeypatch):
    """NBIOT_KERNEL=tpc: every advance runs on the thread-per-cell kernel (csrc/nbiot_simt.cuh: the one-lane build of the core, header
    in local memory, 32 unrelated cells per warp). Same state buffer as the warp-per-cell kernel, byte for byte, on short launches
    (external scheduling agent, 2 carriers), on long ones (NPRACH agent external, single-carrier build) and on run_until; ragged
    batch sizes (not a multiple of the CTA), masked advances."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
    hdr, tk = _cabi.load().nbiot_hdr_bytes(), _cabi.load().nbiot_hdr_task_bytes()

    def snapshot(env, N):
        torch.cuda.synchronize()
        st = env.state.cpu().numpy().copy()
        st[:N * hdr].reshape(N, hdr)[:, hdr - tk:] = 0          # NbHdr.tk_*: only the task-chain kernels write them
        return st
    states = []
    for kern in ('warp', 'tpc', 'tpcv'):
        monkeypatch.setenv('NBIOT_KERNEL', kern)
        out = []
        N = 700
        env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=list(range(N)), n_carriers=2, ue_cap=512, grid_w=2048, hist_cap=64)
        env.reset()
        r = np.random.default_rng(5)
        for _ in range(60):
            env.step(_random_ext_actions(r, env.action_nvec, N))
        env.run_until(6000)
        out.append(snapshot(env, N)); assert env.stats()['faulted'] == 0; env.close()
        N = 333
        env = BatchedNBIoTEnv(dict(CFG1, ratio=0.5), agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, seeds=list(range(N)), ue_cap=1000, grid_w=2048, hist_cap=512)
        env.reset()
        for _ in range(3):
            env.step(_random_ext_actions(r, env.action_nvec, N))
        env.run_until(12000)
        out.append(snapshot(env, N)); assert env.stats()['faulted'] == 0; env.close()
        states.append(out)
    for other in states[1:]:
        assert np.array_equal(states[0][0], other[0]) and np.array_equal(states[0][1], other[1])


def test_per_cell_launch_timing(monkeypatch):
    """NBIOT_PROF_INST=1: start / end time and SM of every cell in the last simulation launch (measurement aid, tools/inst_times.py)"""
    from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
    monkeypatch.setenv('NBIOT_PROF_INST', '1')
    N = 64
    env = BatchedNBIoTEnv(CFG1, num_envs=N, seeds=list(range(N)), ue_cap=256, grid_w=1024)
    env.reset(); env.run_until(3000)
    out = np.zeros((N, 4), dtype=np.uint64)
    _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    assert (out[:, 1] > out[:, 0]).all() and (out[:, 2] < 1024).all()
    env.close()
    monkeypatch.delenv('NBIOT_PROF_INST')
    env = BatchedNBIoTEnv(CFG1, num_envs=2, ue_cap=64, grid_w=256)
    with pytest.raises(Exception):
        _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    env.close()


def test_builds_of_the_core_give_identical_states(monkeypatch):
    """The simulation kernels exist in four builds of the core (csrc/nbiot_simk.cuh): generic / single-carrier x ring anywhere / ring
    staged. A single-carrier batch run on each of them -- short launches (every decision external), long launches (NPRACH agent
    external) and run_until -- ends in the same state buffer, byte for byte."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    N = 200
    states = []
    for variant, staged in ((None, None), ('generic', None), (None, '0'), ('generic', '0')):
        for k, v in (('NBIOT_SIM_VARIANT', variant), ('NBIOT_SIM_STAGED_BUILDS', staged)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        out = []
        env = BatchedNBIoTEnv(dict(CFG1, ratio=0.5), agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, seeds=list(range(N)), ue_cap=512, grid_w=2048, hist_cap=512)
        env.reset()
        r = np.random.default_rng(21)
        for _ in range(3):
            env.step(_random_ext_actions(r, env.action_nvec, N))
        env.run_until(12000)
        torch.cuda.synchronize()
        out.append(env.state.cpu().numpy().copy()); assert env.stats()['faulted'] == 0; env.close()
        env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=list(range(N)), ue_cap=512, grid_w=2048, hist_cap=64)
        env.reset()
        for _ in range(80):
            env.step(_random_ext_actions(r, env.action_nvec, N))
        torch.cuda.synchronize()
        out.append(env.state.cpu().numpy().copy()); assert env.stats()['faulted'] == 0; env.close()
        states.append(out)
    for other in states[1:]:
        assert np.array_equal(states[0][0], other[0]) and np.array_equal(states[0][1], other[1])


def _random_ext_actions(rng, nvec, n):
    return np.stack([rng.integers(0, m, size=n) for m in nvec], axis=1).astype(np.uint8)


@pytest.mark.parametrize('scenario', ['nprach', 'npusch'])
def test_controller_split_against_oracle(scenario, kernel):
    """config 2 / config 3: the external agent acts, the internal fixed-action agents live on the device."""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    if scenario == 'nprach':
        conf, agents, ext, n_steps, N = dict(CFG1, ratio=0.5), NPRACH_AGENTS, 3, 12, 24
    else:
        conf, agents, ext, n_steps, N = CFG3, NPUSCH_AGENTS, 0, 400, 12
    seeds = list(range(100, 100 + N))
    env = BatchedNBIoTEnv(conf, agents_conf=agents, ext_agent=ext, num_envs=N, seeds=seeds, ue_cap=conf['M'], grid_w=2048, hist_cap=1024)
    ctrls = [OracleController(conf, s, agents_conf=agents, ext_agent=ext) for s in seeds]
    obs, infos = env.reset()
    obs = obs.cpu().numpy()
    for i, c in enumerate(ctrls):
        assert np.array_equal(c.reset(), obs[i])
    rng = np.random.default_rng(5)
    nvec = env.action_nvec
    assert env.obs_dim == (26 if scenario == 'nprach' else 57)
    for step in range(n_steps):
        a = _random_ext_actions(rng, nvec, N)
        if scenario == 'nprach':
            a[:, 0] = np.minimum(a[:, 0], a[:, 1])      # keep th_C1 <= th_C0 most of the time; both orders are legal
        if step % 2 == 0:
            obs, rew, term, trunc, infos = env.step(a)
            obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
            assert not term.any().item() and not trunc.any().item()
        else:
            obs, rew = env.step_host(a)                 # the host-buffer entry point must agree with the device one
        recs = env.records()
        for i, c in enumerate(ctrls):
            o_obs, o_rew = c.step(a[i])
            assert np.array_equal(o_obs, obs[i]), (step, i)
            assert o_rew == rew[i]
            assert not record_diff(c.rec, recs[i], rtol=0.0)
        if scenario == 'nprach':
            assert (recs['time'] == 3000 * (step + 1)).all()
            info = infos[0] if step % 2 == 0 else None
            if info is not None:
                inc, dl, sv = ctrls[0].hist
                assert info['delays'] == dl.tolist() and info['service_times'] == sv.tolist() and np.allclose(info['incoming'], inc)
    env.close()


def test_run_until_on_internal_agents(kernel):
    """config 1 / 5 style: no external agent, default agents on the device, one launch to the horizon."""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    N, T = 16, 30000
    seeds = list(range(N))
    env = BatchedNBIoTEnv(CFG1, num_envs=N, seeds=seeds, ue_cap=512, grid_w=2048)
    env.reset()
    env.run_until(T)
    recs = env.records()
    st = env.stats()
    tot = np.zeros(3, dtype=np.int64)
    for i, s in enumerate(seeds):
        c = OracleController(CFG1, s)
        c.reset()
        c.run_until(T)
        assert not record_diff(c.rec, recs[i], rtol=0.0), i
        tot += np.asarray(c.cell.counters()['ue_departures'])
    assert st['ue_departures'] == tot.tolist()
    assert st['subframes'] == int(recs['time'].sum()) and st['faulted'] == 0
    # a second call with the same horizon does nothing; a later horizon continues
    env.run_until(T)
    assert np.array_equal(env.records()['time'], recs['time'])
    env.close()


def test_run_until_with_an_external_agent_registered(kernel):
    """Controller.run_until / single_step (controller.py:156-195) let EVERY agent of a state act -- the one registered as external
    with its fixed action -- and run to the horizon; they do not stop at the external agent's first decision."""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    for conf, agents, ext, T in ((dict(CFG1, ratio=0.5), NPRACH_AGENTS, 3, 9000), (CFG3, NPUSCH_AGENTS, 0, 4000)):
        N = 6
        seeds = list(range(40, 40 + N))
        env = BatchedNBIoTEnv(conf, agents_conf=agents, ext_agent=ext, num_envs=N, seeds=seeds, ue_cap=conf['M'], grid_w=2048, hist_cap=1024)
        env.reset()
        env.run_until(T)
        recs = env.records()
        assert (recs['time'] >= T).all()
        for i, s in enumerate(seeds):
            c = OracleController(conf, s, agents_conf=agents, ext_agent=ext)
            c.reset()
            c.run_until(T)
            assert not record_diff(c.rec, recs[i], rtol=0.0), i
        # and the external agent takes over again afterwards
        a = _random_ext_actions(np.random.default_rng(1), env.action_nvec, N)
        env.step(a)
        assert (env.records()['time'] > recs['time']).all()
        env.close()


def test_batch_independence_and_determinism_at_scale():
    """4096 cells (config 2 size): instance i of a large batch equals the same seed simulated alone,
    two runs are identical, and the reduced statistics equal the sum of the per-instance records."""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    N = 4096
    conf = dict(CFG1, ratio=0.5)
    rng = np.random.default_rng(11)
    acts = [_random_ext_actions(rng, [15, 15, 4, 4, 4, 6, 6, 6], N) for _ in range(3)]
    out = []
    for rep in range(2):
        env = BatchedNBIoTEnv(conf, agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, ue_cap=512, grid_w=2048)
        env.reset()
        for a in acts:
            obs, rew, _, _, infos = env.step(a)
        out.append((obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), env.records(), env.stats()))
        env.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert out[0][2].tobytes() == out[1][2].tobytes()
    recs, st = out[0][2], out[0][3]
    assert not (recs['fault'] & 0x1F).any() and (recs['time'] == 9000).all()
    assert st['subframes'] == 9000 * N
    # every one of the 4096 cells against the CPU oracle (threaded C driver), bit for bit
    from oracle import oracle_py
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.record import make_conf
    orecs = oracle_py.run_batch(make_conf(conf, ue_cap=512, grid_w=2048), ControllerTables(NPRACH_AGENTS, 3).build(), np.arange(N),
                                actions=np.stack(acts), steps=len(acts))
    bad = [i for i in range(N) if record_diff(orecs[i], recs[i], rtol=0.0)]
    assert not bad, f'{len(bad)} of {N} cells differ from the oracle, first: {bad[:5]}'
    for i in (0, 1, 777, 4095):
        c = OracleController(conf, i, agents_conf=NPRACH_AGENTS, ext_agent=3)
        c.reset()
        for a in acts:
            o, r = c.step(a[i])
        assert np.array_equal(o, out[0][0][i]) and r == out[0][1][i]
        assert not record_diff(c.rec, recs[i], rtol=0.0)


def test_faults_are_never_silent():
    from nb_iot_environment_b200 import BatchedNBIoTEnv, parameters as par
    # too few UE slots for the offered load: the instance stops and says so
    env = BatchedNBIoTEnv({'M': 20000, 'ratio': 0.2, 'buffer_range': [256, 256], 'reward_criteria': 'throughput'}, num_envs=4, ue_cap=64)
    env.reset()
    env.run_until(20000)
    recs, st = env.records(), env.stats()
    assert (recs['fault'] & 1).all() and st['faulted'] == 4 and (recs['time'] < 20000).all()
    env.close()
    # an action the reference raises on (sc = 2 -> N_sc = 9, KeyError)
    env = BatchedNBIoTEnv(CFG1, num_envs=2, all_states_external=True)
    env.reset()
    a = np.tile(np.asarray(par.default_action_vector(), dtype=np.uint8), (2, 1))
    bad = a.copy()
    bad[1, par.control_items['sc']] = 2
    for _ in range(200):
        env.node_step(bad)
    recs = env.records()
    assert recs['fault'][0] == 0 and recs['fault'][1] & 4
    env.close()


def test_checkpoint_resume_and_ragged_batch():
    """state snapshot -> restore in a fresh environment continues bit-identically; batch sizes that do not fill a CTA"""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    for N in (1, 5):
        seeds = list(range(200, 200 + N))
        env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=seeds, ue_cap=1024, grid_w=2048)
        env.reset()
        rng = np.random.default_rng(3)
        acts = [_random_ext_actions(rng, env.action_nvec, N) for _ in range(60)]
        for a in acts[:30]:
            env.step(a)
        snap = env.save_state()
        for a in acts[30:]:
            obs_a, rew_a, _, _, _ = env.step(a)
        rec_a, obs_a, rew_a = env.records(), obs_a.cpu().numpy().copy(), rew_a.cpu().numpy().copy()
        env.close()
        env2 = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=[0] * N, ue_cap=1024, grid_w=2048)
        env2.load_state(snap)
        for a in acts[30:]:
            obs_b, rew_b, _, _, _ = env2.step(a)
        assert env2.records().tobytes() == rec_a.tobytes()
        assert np.array_equal(obs_b.cpu().numpy(), obs_a) and np.array_equal(rew_b.cpu().numpy(), rew_a)
        env2.close()
        c = OracleController(CFG3, seeds[-1], agents_conf=NPUSCH_AGENTS, ext_agent=0)
        c.reset()
        for a in acts:
            c.step(a[-1])
        assert not record_diff(c.rec, rec_a[-1], rtol=0.0)


def test_reset_after_run_and_agent_reconfiguration():
    """the scripts' sequence Controller(...).reset(); set_ext_agent(k); env.reset() (evaluate_RL.py:128-131), a second
    reset after running (NodeB.reset keeps total_departures / NPDCCH_sf_left), and
    DummyAgent.set_action on an internal agent"""
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    N = 6
    seeds = list(range(300, 300 + N))
    conf = dict(CFG1, ratio=0.5)
    env = BatchedNBIoTEnv(conf, agents_conf=NPRACH_AGENTS, ext_agent=-1, num_envs=N, seeds=seeds, ue_cap=1000, grid_w=2048)
    ctrls = [OracleController(conf, s, agents_conf=NPRACH_AGENTS, ext_agent=-1) for s in seeds]
    env.reset()
    env.set_ext_agent(3)
    env.set_internal_action(2, {'backoff': 5, 'mac_timer': 4})
    obs, _ = env.reset()
    for c in ctrls:
        c.reset()
        c.t.set_ext_agent(3)
        c.t.agents[2].set_action({'backoff': 5, 'mac_timer': 4})
        o = c.reset()
    rng = np.random.default_rng(9)
    for rnd in range(2):
        for _ in range(4):
            a = _random_ext_actions(rng, env.action_nvec, N)
            a[:, 0] = np.minimum(a[:, 0], a[:, 1])
            obs, rew, _, _, _ = env.step(a)
            recs = env.records()
            for i, c in enumerate(ctrls):
                o, r = c.step(a[i])
                assert np.array_equal(o, obs[i].cpu().numpy()) and r == rew[i].item()
                assert not record_diff(c.rec, recs[i], rtol=0.0)
        obs, _ = env.reset()            # reset in the middle of an episode
        recs = env.records()
        for i, c in enumerate(ctrls):
            o = c.reset()
            assert np.array_equal(o, obs[i].cpu().numpy())
            assert not record_diff(c.rec, recs[i], rtol=0.0)
    env.close()


def test_api_guards():
    """Round-1 review items: an infos object of an earlier step refuses to serve later data; load_state refreshes the cached
    observation; action values that a uint8 cast would wrap are refused; the all-external action space has one entry per index of
    the shared vector; a later, smaller batch does not lower the shared-memory limit of an earlier one."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi, parameters as par
    N = 4
    env = BatchedNBIoTEnv(dict(CFG1, ratio=0.5), agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, n_carriers=2, ue_cap=1000, grid_w=4096, hist_cap=512)
    small = BatchedNBIoTEnv(CFG1, num_envs=2, ue_cap=64, grid_w=256)           # needs far less shared memory per CTA than `env`
    small.reset(); small.run_until(500)
    obs0, infos0 = env.reset()
    assert infos0[0]['state'] == 'NPRACH_update'                               # read in time: fine, and cached
    a = _random_ext_actions(np.random.default_rng(2), env.action_nvec, N)
    obs1, rew1, _, _, infos1 = env.step(a, fresh=True)
    keep1 = obs1.clone()
    snap = env.save_state()
    obs2, rew2, _, _, infos2 = env.step(a, validate=True)
    assert infos0[0]['time'] == 0                                              # served from its own cache
    with pytest.raises(_cabi.NbiotError):
        infos1[0]                                                              # never read before the environment moved on
    assert infos2[0]['time'] == 6000
    assert torch.equal(obs1, keep1) and not torch.equal(obs2, keep1)           # fresh=True tensors are not overwritten by the next step
    env.load_state(snap)
    assert torch.equal(env._obs[:, :env.obs_dim], keep1)                       # the cached observation follows the restored records
    with pytest.raises(ValueError):
        env.step(a.astype(np.int64) + 256)
    with pytest.raises(ValueError):
        env.step(-a.astype(np.int64) - 1)
    bad = a.copy(); bad[0, 2] = 200
    with pytest.raises(ValueError):
        env.step(bad, validate=True)
    env.close(); small.close()
    ext_all = BatchedNBIoTEnv(CFG1, num_envs=1, all_states_external=True, ue_cap=64, grid_w=256)
    nvec = ext_all.action_nvec
    assert len(nvec) == par.N_actions == 23 and nvec[par.control_items['Imcs']] == 7 and nvec[par.control_items['id']] == 4
    ext_all.close()


def _boundary_cells(N, per_cell_bytes):
    """cells on both sides of every 4 GiB multiple inside an instance-major array of `per_cell_bytes` per cell"""
    out = set()
    k = 1
    while k * (1 << 32) < N * per_cell_bytes:
        i = (k * (1 << 32)) // per_cell_bytes
        out.update(j for j in (i - 1, i, i + 1) if 0 <= j < N)
        k += 1
    return out


def _sample_cells(N, arrays, n_random=160, seed=7):
    idx = set(range(16)) | set(range(N - 16, N))
    for b in arrays:
        idx |= _boundary_cells(N, b)
    rng = np.random.default_rng(seed)
    idx |= set(int(i) for i in rng.integers(0, N, size=n_random))
    return np.asarray(sorted(idx), dtype=np.int64)


def test_parity_beyond_4gib_config3_setup():
    """BASELINE configs[2] set-up (NPUSCH scenario, scheduling agent external, one launch per decision) with 131 072 cells: 21.7 GB
    of state, every per-cell array several times past 4 GiB. 250+ sampled cells -- the first and last ones, both sides of every
    4 GiB boundary of every array, random ones -- against the CPU oracle, bit for bit."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.record import make_conf, RECORD_DTYPE
    from oracle import oracle_py
    N, STEPS, caps = 131072, 40, dict(ue_cap=1024, grid_w=2048, hist_cap=64)
    env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, **caps)
    assert env.state.numel() > 4 * (1 << 32)
    ra_cap = caps['ue_cap'] + 64
    arrays = [3504, 2 * caps['grid_w'], 3 * ra_cap * 16, caps['ue_cap'] * 16, caps['ue_cap'] * 64, caps['ue_cap'] * 2, RECORD_DTYPE.itemsize, ra_cap * 2]
    idx = _sample_cells(N, arrays, n_random=260)
    assert len(idx) >= 256
    g = torch.Generator(device='cuda'); g.manual_seed(11)
    nvec = torch.tensor(env.action_nvec, device='cuda')
    env.reset()
    acts = []
    for s in range(STEPS):
        a = (torch.rand((N, 3), device='cuda', generator=g) * nvec).to(torch.uint8)
        env.step(a)
        acts.append(a[torch.as_tensor(idx, device='cuda')].cpu().numpy())
    recs = env.records()
    assert env.stats()['faulted'] == 0
    orecs = oracle_py.run_batch(make_conf(CFG3, **caps), ControllerTables(NPUSCH_AGENTS, 0).build(), idx.astype(np.uint64), actions=np.stack(acts), steps=STEPS)
    bad = [int(i) for k, i in enumerate(idx) if record_diff(orecs[k], recs[i], rtol=0.0)]
    assert not bad, f'{len(bad)} of {len(idx)} sampled cells differ from the oracle, first: {bad[:5]}'
    env.close()


def test_parity_beyond_4gib_two_carriers():
    """BASELINE configs[3] set-up (two carriers, M=3000 bursty, every decision external) with 24 576 cells x 447 KB = 11 GB of state:
    sampled cells on both sides of every 4 GiB boundary against the CPU oracle, bit for bit."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.record import make_conf, RECORD_DTYPE
    from oracle import oracle_py
    CFG4 = {'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True}
    N, STEPS, caps = 24576, 120, dict(ue_cap=3000, grid_w=2048, hist_cap=64)
    env = BatchedNBIoTEnv(CFG4, num_envs=N, n_carriers=2, all_states_external=True, **caps)
    assert env.state.numel() > 2 * (1 << 32)
    ra_cap = caps['ue_cap'] + 64
    idx = _sample_cells(N, [2 * 2 * caps['grid_w'], 3 * ra_cap * 16, caps['ue_cap'] * 16, caps['ue_cap'] * 64, caps['ue_cap'] * 2], n_random=260)
    mx = torch.tensor([1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, 7, 5, 7, 7, 5, 7, 7, 14, 14], device='cuda') + 1
    g = torch.Generator(device='cuda'); g.manual_seed(12)
    env.reset()
    acts = []
    sel = torch.as_tensor(idx, device='cuda')
    for s in range(STEPS):
        a = (torch.rand((N, 23), device='cuda', generator=g) * mx).to(torch.uint8)
        a[:, 5] = torch.where(a[:, 5] == 2, torch.full_like(a[:, 5], 3), a[:, 5])          # sc = 2 is invalid
        rar = env.record_field('state') == 2                                              # RAR_window: ce_level / rar_Imcs share items 1 / 2
        a[:, 1] = torch.where(rar, a[:, 1] % 3, a[:, 1]); a[:, 2] = torch.where(rar, a[:, 2] % 3, a[:, 2])
        env.node_step(a)
        acts.append(a[sel].cpu().numpy())
    recs = env.records()
    assert env.stats()['faulted'] == 0
    orecs = oracle_py.run_batch(make_conf(CFG4, n_carriers=2, **caps), ControllerTables(n_carriers=2).build(all_states_external=True), idx.astype(np.uint64),
                                actions=np.stack(acts), steps=STEPS)
    bad = [int(i) for k, i in enumerate(idx) if record_diff(orecs[k], recs[i], rtol=0.0)]
    assert not bad, f'{len(bad)} of {len(idx)} sampled cells differ from the oracle, first: {bad[:5]}'
    env.close()


def test_info_lists_longer_than_the_record():
    """info['receptions'] / ['errors'] / ['access'] accumulate between two Scheduling decisions (perf_monitor.py:92-103, cleared at
    node_b.py:589); in a loaded two-carrier cell that stays in RAR windows, info['access'] gains an entry per connection and outgrows the
    record's 16 entries by far (325 in 600 decisions of the oracle). The rest lives in the side buffer: infos returns the complete
    lists, equal to the oracle's, and nothing is flagged as truncated."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle.oracle_py import OracleCell
    CFG4 = {'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True}
    N, STEPS = 1024, 700
    env = BatchedNBIoTEnv(CFG4, num_envs=N, n_carriers=2, all_states_external=True, ue_cap=3000, grid_w=2048, hist_cap=2048)
    mx = torch.tensor([1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, 7, 5, 7, 7, 5, 7, 7, 14, 14], device='cuda') + 1
    g = torch.Generator(device='cuda'); g.manual_seed(12)
    env.reset()
    acts, found = [], None
    for s in range(STEPS):
        a = (torch.rand((N, 23), device='cuda', generator=g) * mx).to(torch.uint8)
        a[:, 5] = torch.where(a[:, 5] == 2, torch.full_like(a[:, 5], 3), a[:, 5])
        rar = env.record_field('state') == 2
        a[:, 1] = torch.where(rar, a[:, 1] % 3, a[:, 1]); a[:, 2] = torch.where(rar, a[:, 2] % 3, a[:, 2])
        env.node_step(a)
        acts.append(a.cpu().numpy())
        if s > 300 and s % 20 == 0:
            long_ = (env.record_field('n_acc') > 80) & (env.record_field('state') == 0)
            if bool(long_.any()):
                found = int(torch.nonzero(long_)[0])
                break
    assert found is not None, 'no cell accumulated a long access list: make the scenario heavier'
    from nb_iot_environment_b200.env import LazyInfos
    info = LazyInfos(env)[found]
    assert len(info['access']) == int(env.records()[found]['n_acc']) > 80
    cell = OracleCell(CFG4, found, n_carriers=2)
    cell.reset()
    for a in acts:
        cell.step(a[found])
    rx, err, unfit, acc = cell.step_lists()
    assert info['receptions'] == rx and info['errors'] == err and info['unfit'] == unfit and info['access'] == acc
    assert env.stats()['truncated'] == 0 and env.stats()['faulted'] == 0
    env.close()


def test_dealing_cells_to_sms_changes_nothing(monkeypatch):
    """One-wave long launches of the warp-per-cell kernel deal the cells to the SMs by predicted cost (nbiot_deal_kernel: the cell's previous
    duration + the length of its random-access lists, sorted and dealt in snake order). Which warp of which SM runs a cell never changes
    a result: same state buffer with the dealing off, masked advances included."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    N = 1000
    states = []
    for deal in ('1', '0'):
        monkeypatch.setenv('NBIOT_DEAL', deal)
        monkeypatch.setenv('NBIOT_KERNEL', 'warp')
        env = BatchedNBIoTEnv(dict(CFG1, ratio=0.5), agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, seeds=list(range(N)), ue_cap=1000, grid_w=2048, hist_cap=512)
        env.reset()
        r = np.random.default_rng(33)
        for k in range(5):
            a = _random_ext_actions(r, env.action_nvec, N)
            if k == 3:
                env.step(a, active=torch.as_tensor(r.integers(0, 2, size=N).astype(bool)))
            else:
                env.step(a)
        env.run_until(18000)
        torch.cuda.synchronize()
        states.append(env.state.cpu().numpy().copy()); assert env.stats()['faulted'] == 0
        assert (env.records()['time'] >= 15000).all()
        env.close()
    assert np.array_equal(states[0], states[1])


def test_kernel_selector_of_the_constructor(monkeypatch):
    """BatchedNBIoTEnv(kernel=...) / nbiot_conf_t.kernel: which kernel runs the long advances; the state buffers do not depend on it"""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    monkeypatch.delenv('NBIOT_KERNEL', raising=False)
    states, names = [], []
    for kernel in (None, 'warp', 'thread', 'thread_voted'):
        env = BatchedNBIoTEnv(CFG1, num_envs=70, seeds=list(range(70)), ue_cap=512, grid_w=2048, kernel=kernel)
        env.reset()
        env.run_until(12000)
        torch.cuda.synchronize()
        names.append(env.kernel_name())
        hdr = env.L.nbiot_hdr_bytes(); tail = env.L.nbiot_hdr_task_bytes()
        s = env.state.cpu().numpy().copy()
        h = s[:70 * hdr].reshape(70, hdr); h[:, hdr - tail:] = 0      # the trailing task words only the task-chain kernels write
        states.append(s); assert env.stats()['faulted'] == 0
        env.close()
    assert 'warp' in names[0] and 'warp' in names[1] and 'thread' in names[2] and 'voted' in names[3], names
    for s in states[1:]:
        assert np.array_equal(states[0], s)


def test_automatic_kernel_choice_follows_the_load(monkeypatch):
    """kernel=None on a batch large enough for the thread-per-cell kernel: the mean number of live UEs per cell, read back behind every long
    launch on the device, decides which of the two kernels runs it -- lightly loaded cells stay on the voted thread-per-cell kernel, cells with long lists go to
    the warp-per-cell kernel (profiles/r02_tpc_variants.txt section 16). The choice never changes a result (sampled cells against the oracle)."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from oracle_controller import OracleController
    monkeypatch.delenv('NBIOT_KERNEL', raising=False)
    N = 32768
    # light: configs[0] traffic on the default agents
    env = BatchedNBIoTEnv(CFG1, num_envs=N, seeds=list(range(N)), ue_cap=256, grid_w=1024, hist_cap=64)
    env.reset()
    for T in (3000, 6000, 9000):
        env.run_until(T)
        torch.cuda.synchronize()
    assert 'voted' in env.kernel_name() and env.stats()['faulted'] == 0
    env.close()
    # loaded: the NPRACH scenario of bench.py (bursty traffic, 100+ live UEs per cell after the first bursts)
    conf = dict(CFG1, ratio=0.5)
    env = BatchedNBIoTEnv(conf, agents_conf=NPRACH_AGENTS, ext_agent=3, num_envs=N, seeds=list(range(N)), ue_cap=1000, grid_w=2048, hist_cap=512)
    env.reset()
    r = np.random.default_rng(11)
    a = _random_ext_actions(r, env.action_nvec, 1)
    a[:, 0] = np.minimum(a[:, 0], a[:, 1])
    acts = np.repeat(a, N, axis=0)
    names = []
    for k in range(6):
        env.step(acts)
        torch.cuda.synchronize()
        names.append(env.kernel_name())
    st = env.stats()
    assert st['faulted'] == 0 and st['live_ues'] / N > 140, st        # NB_TPC_MAX_LOAD
    assert 'voted' in names[0] and any('warp' in n for n in names[1:]), names
    recs = env.records()
    for i in (0, 1, N // 2, N - 1):
        c = OracleController(conf, i, agents_conf=NPRACH_AGENTS, ext_agent=3)
        c.reset()
        for k in range(6):
            c.step(acts[i])
        assert not record_diff(c.rec, recs[i], rtol=0.0), i
    env.close()
