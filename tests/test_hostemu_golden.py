"""The device simulation core (csrc/nbiot_core.cuh), compiled for the CPU with emulated lanes, against the
golden traces of the Python reference. This is a debugging aid that exercises the exact kernel source
without a GPU: 1 lane (scalar logic) on every case, 32 fiber lanes (ballot / shuffle / rank logic) on a
shorter prefix. The GPU itself is checked in test_gpu_parity.py."""
import numpy as np
import pytest

from golden_util import CASES, Golden, record_diff, assert_pm_lists, REF_RTOL

CAPS = {'capture_effect': dict(grid_w=2048), 'capture_effect_two_carriers': dict(grid_w=2048), 'traces_two_carriers': dict(grid_w=2048), 'manual_branches': dict(grid_w=4096), 'cfg3_npusch_random_policy': dict(grid_w=2048), 'cfg2_random_policy': dict(grid_w=2048),
        'cfg4_two_carriers_bursty': dict(grid_w=2048)}


def _replay(case, lanes, decisions):
    from hostemu_py import EmuBatch
    g = Golden(case)
    caps = dict(ue_cap=g.conf['M'], grid_w=1024, hist_cap=1024)
    caps.update(CAPS.get(case, {}))
    b = EmuBatch(g.conf, g.seeds, n_carriers=g.n_carriers, lanes=lanes, capture_effect=g.capture_effect, **caps)
    recs = b.reset()
    D = min(decisions, g.decisions)
    for d in range(-1, D):
        if d >= 0:
            recs, steps = b.advance(g.actions[:, d, :].astype(np.uint8), max_steps=1)
            assert (steps == 1).all()
        for si in range(len(g.seeds)):
            diff = record_diff(g.records[si, d + 1], recs[si], rtol=REF_RTOL)
            assert not diff, f'{case} seed {g.seeds[si]} decision {d + 1}: {diff[:3]}'
    assert not (recs['fault'] & 0x1F).any()
    if D == g.decisions:
        for si in range(len(g.seeds)):
            c, gc = b.counters(si), g.counters[si]
            for k, v in gc.items():
                if k == 'av_delay':
                    assert c[k] == pytest.approx(v, rel=REF_RTOL)
                else:
                    assert c[k] == v, k
            gl = g.pm_lists(si)
            if gl is not None:                                   # traces / statistics sample lists (perf_monitor.py:61-79)
                from nb_iot_environment_b200.record import samples_to_lists
                rec_, n_ = b.samples(si)
                assert n_ == len(rec_)
                assert_pm_lists(samples_to_lists(rec_), gl, f'{case} seed {g.seeds[si]}')
            gs = g.stats(si)
            if gs is not None:                                   # statistics histories (perf_monitor.py:161-171)
                st = b.stats(si)
                assert st['n'] == len(gs['delay'])
                for k in ('delay', 'connection', 'access'):
                    assert np.array_equal(st[k], gs[k]), k
                assert np.array_equal(st['bits'].astype(np.float64) / st['delay'].astype(np.float64), gs['th'])


@pytest.mark.parametrize('case', CASES)
def test_core_scalar_logic(case, oracle_lib):
    _replay(case, 1, 10 ** 9)


@pytest.mark.parametrize('case', ['cfg1_default', 'cfg4_two_carriers_bursty', 'overload', 'capture_effect'])
def test_core_lane_parallel_sections(case, oracle_lib):
    _replay(case, 32, 400)


@pytest.mark.parametrize('case', CASES)
def test_core_task_chain(case, oracle_lib):
    """The same core cut into tasks (nbiot_core.cuh "tasks"), the instance stored and reloaded between two tasks:
    the control flow of the migrating-instance kernel, against the same golden traces."""
    _replay(case, '1t', 10 ** 9)


@pytest.mark.parametrize('case', CASES)
def test_core_task_chain_fine_grained(case, oracle_lib):
    """The task chain as the voted-dispatch thread-per-cell kernel cuts it (csrc/nbiot_simt.cuh): the glue between two handlers is a
    task class of its own and handles ONE event per task, so that the lanes of a warp pop their events in lock step."""
    _replay(case, '1tg', 10 ** 9)


@pytest.mark.parametrize('case', CASES)
def test_core_with_block_summaries(case, oracle_lib):
    """The one-lane core with the thread-per-cell kernel's block summaries of the look-ahead ring (window scans read one summary per
    whole 32-column block): rebuilt from the ring at every decision here, maintained by every fill in between."""
    _replay(case, '1b', 10 ** 9)


@pytest.mark.parametrize('case', ['cfg1_default', 'cfg3_npusch_random_policy'])
def test_event_selection_second_pass(case, oracle_lib):
    """The event selection looks at 32 keys per pass and takes a second pass only when pool entries above the first 22 are in use
    (more than 22 pending events). This build hands the pool out from its highest entry, so every trace goes through the second pass."""
    _replay(case, '32h', 300)


@pytest.mark.parametrize('lanes', [1, '1tg', '1b'])
def test_run_until_with_an_external_agent_registered(lanes, oracle_lib):
    """Controller.run_until (controller.py:156-195) with an external agent registered: every agent of a state acts with its fixed
    action and the cells run to the horizon (the core used to hand control back at the external agent's first decision)."""
    from hostemu_py import EmuBatch
    from nb_iot_environment_b200.controller import ControllerTables
    from oracle_controller import OracleController
    from test_gpu_parity import NPRACH_AGENTS
    conf = {'ratio': 0.5, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
    seeds = [3, 4]
    b = EmuBatch(conf, seeds, lanes=lanes, ctrl=ControllerTables(NPRACH_AGENTS, 3).build(), ue_cap=1000, grid_w=2048, hist_cap=1024)
    b.reset()
    T = 20000 if lanes == '1b' else 7000          # one long advance: the block summaries are built once and maintained for thousands of fills
    recs, steps = b.advance(None, until_time=T)
    assert (recs['time'] >= T).all() and (steps > 100).all()
    for i, s in enumerate(seeds):
        c = OracleController(conf, s, agents_conf=NPRACH_AGENTS, ext_agent=3)
        c.reset()
        c.run_until(T)
        assert not record_diff(c.rec, recs[i], rtol=0.0)
