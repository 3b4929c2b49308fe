"""The CPU oracle (oracle/nbiot_oracle.c) against the golden traces of the unmodified Python
reference: integer fields bit-exact, float fields to 1e-9 relative, same number of random draws."""
import numpy as np
import pytest

from golden_util import CASES, Golden, record_diff, assert_pm_lists, REF_RTOL


@pytest.mark.parametrize('case', CASES)
def test_oracle_replays_reference_trace(case, oracle_lib):
    from oracle.oracle_py import OracleCell
    g = Golden(case)
    for si, seed in enumerate(g.seeds):
        cell = OracleCell(g.conf, seed, n_carriers=g.n_carriers, capture_effect=g.capture_effect)
        cell.enable_event_log()
        rec = cell.reset()
        assert not record_diff(g.records[si, 0], rec, rtol=REF_RTOL)
        hists = g.hist(si)
        hi = 1 if g.records[si, 0]['state'] == 1 else 0      # the reset record is an NPRACH_update flavour
        for d in range(g.decisions):
            rec = cell.step(g.actions[si, d])
            diff = record_diff(g.records[si, d + 1], rec, rtol=REF_RTOL)
            assert not diff, f'{case} seed {seed} decision {d + 1}: {diff[:4]}'
            if rec['state'] == 1:
                inc, dl, sv = cell.hist()
                ginc, gdl, gsv = hists[hi]
                hi += 1
                assert np.array_equal(dl, gdl) and np.array_equal(sv, gsv)
                assert np.allclose(inc, ginc, rtol=REF_RTOL, atol=0)
        c, gc = cell.counters(), g.counters[si]
        for k, v in gc.items():
            if k == 'av_delay':
                assert c[k] == pytest.approx(v, rel=REF_RTOL)
            else:
                assert c[k] == v, k
        gs = g.stats(si)
        if gs is not None:                                       # statistics=True: the per-departure histories (perf_monitor.py:161-171)
            st = cell.stats()
            for k in ('delay', 'connection', 'access'):
                assert np.array_equal(st[k], gs[k]), k
            assert np.array_equal(st['th'], gs['th'])           # acked bits / delay: one IEEE division of the same integers
            assert len(gs['delay']) == sum(gc['ue_departures'])
        gl = g.pm_lists(si)
        if gl is not None:                                       # traces / statistics: PerfMonitor's sample lists (perf_monitor.py:61-79)
            from nb_iot_environment_b200.record import samples_to_lists
            assert_pm_lists(samples_to_lists(cell.samples()), gl, f'{case} seed {seed}')
        ev = g.events(si)
        assert np.array_equal(cell.event_log()[:len(ev)], ev)   # same pops in the same (time, seq) order
