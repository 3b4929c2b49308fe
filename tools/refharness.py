#!/usr/bin/env python3
"""Drive the UNMODIFIED reference (/root/reference) in the build container.

Used only to generate golden vectors (tools/gen_golden.py) and for ad-hoc comparisons; the
GPU box has no /root/reference, so nothing that runs there imports this module.

 * matplotlib / gymnasium are absent from the image: import-only stand-ins live in
   tools/ref_stubs (the reference only *uses* them in animation code and in wrappers).
 * The reference takes its random generator as an argument (system/system_creator.py:22);
   `PhiloxShim` implements the calls it makes with the law of
   include/nbiot_rng.h, evaluated by the compiled host twins so that the arithmetic is the
   same compiled code the CUDA kernels and the CPU oracle use.
 * The carrier count is a module global read at import time (system/parameters.py:16,96-100;
   system/carrier.py:55,418), so one process can only ever run one carrier count:
   call `load_reference(n_carriers)` once, first.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('NBIOT_REFERENCE', '/root/reference')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from nb_iot_environment_b200.record import RECORD_DTYPE, STATE_IDS, NPRACH_ITEMS, MAX_OCC, STEP_LIST  # noqa: E402
from oracle import oracle_py  # noqa: E402  (host twins of the random stream)

_loaded = None


def load_reference(n_carriers=1):
    """Import the reference with the requested carrier count; returns the modules used here."""
    global _loaded
    if _loaded is not None:
        if _loaded['n_carriers'] != n_carriers:
            raise RuntimeError('the reference fixes its carrier count at import time: use a fresh process')
        return _loaded
    sys.path.insert(0, os.path.join(HERE, 'ref_stubs'))
    sys.path.insert(0, os.path.join(REF, 'gym-system'))
    sys.path.insert(0, REF)
    import system.parameters as par
    par.set_global_parameters(Nc=n_carriers)
    from system.system_creator import create_system
    import system.event_manager as em
    import controller as ref_controller
    import control_agents as ref_agents
    _loaded = {'n_carriers': n_carriers, 'par': par, 'create_system': create_system, 'em': em,
               'controller': ref_controller, 'agents': ref_agents}
    return _loaded


class PhiloxShim:
    """Duck-typed numpy Generator: one Philox counter per scalar draw (include/nbiot_rng.h)."""

    def __init__(self, seed, stream=0):
        self.L = oracle_py.lib()
        self.seed = int(seed)
        self.stream = int(stream)
        self.ctr = 0
        self._pair = (ctypes.c_double * 2)()

    def _next(self):
        c = self.ctr
        self.ctr += 1
        return c

    def random(self, size=None):
        if size is None:
            return self.L.orc_rng_u01(self.seed, self._next(), self.stream)
        if size == 2:
            self.L.orc_rng_pair(self.seed, self._next(), self.stream, self._pair)
            return [self._pair[0], self._pair[1]]
        raise NotImplementedError('random(size) is only used with size=2 (system/channel.py:112)')

    def uniform(self, a=0.0, b=1.0):
        return a + (b - a) * self.L.orc_rng_u01(self.seed, self._next(), self.stream)

    def normal(self, mu=0.0, sigma=1.0):
        return mu + sigma * self.L.orc_rng_normal(self.seed, self._next(), self.stream)

    def binomial(self, n, p):
        s = 0
        for _ in range(int(n)):
            s += 1 if self.L.orc_rng_u01(self.seed, self._next(), self.stream) < p else 0
        return s

    def integers(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        low, high = int(low), int(high)
        if size is None:
            return low + int(self.L.orc_rng_bounded(self.seed, self._next(), self.stream, high - low))
        return [low + int(self.L.orc_rng_bounded(self.seed, self._next(), self.stream, high - low)) for _ in range(int(size))]


def info_to_record(info, reward, rng=None, hist_out=None):
    """reference info dict (node_b.py:299-408) -> one RECORD_DTYPE record."""
    r = np.zeros((), dtype=RECORD_DTYPE)
    r['reward'] = float(reward)
    r['time'] = info['time']
    r['state'] = STATE_IDS[info['state']]
    r['total_ues'] = info['total_ues']
    busy = np.rint((1.0 - np.asarray(info['carrier_state'], dtype=np.float64)) * 12.0)
    r['carrier_busy'] = busy.astype(np.uint8)
    st = info['state']
    if st in ('Scheduling', 'RAR_window_end'):
        n = len(info['ues'])
        r['n_sel'] = n
        r['ues'][:n] = info['ues']
        r['connection_time'] = info['connection_time']
        r['loss'] = info['loss']
        r['sinr'] = info['sinr']
        r['buffer'] = info['buffer']
        for cnt, ids, vals, src in (('n_rx', 'rx_id', 'rx_delay', info['receptions']), ('n_acc', 'acc_id', 'acc_delay', info['access'])):
            r[cnt] = len(src)
            for i, (k, v) in enumerate(list(src.items())[:STEP_LIST]):
                r[ids][i] = k
                r[vals][i] = v
        for cnt, ids, src in (('n_err', 'err_id', info['errors']), ('n_unfit', 'unfit_id', info['unfit'])):
            r[cnt] = len(src)
            r[ids][:min(len(src), STEP_LIST)] = src[:STEP_LIST]
    elif st == 'RAR_window':
        for key in ('ues_per_CE', 'RAR_in', 'RAR_sent', 'RAR_ids', 'RAR_detected', 'RAR_failed', 'action'):
            r[key] = info[key]
        r['valid_CE_control'] = int(info['valid_CE_control'])
        r['NPDCCH_sf_left'] = info['NPDCCH_sf_left']
        r['total_departures'] = info['total_departures']
        # the RAR flavour only exposes 8 of the 16 items; the record carries the full conf, so take it from the node
        r['nprach_conf'] = [info['_full_conf'][k] for k in NPRACH_ITEMS]
    else:
        for key in ('detection_ratios', 'colision_ratios', 'msg3_detection', 'msg3_sent', 'msg3_received', 'distribution',
                    'preambles', 'RA_occupation', 'NPUSCH_occupation', 'NPRACH_occupation', 'av_delay', 'departures'):
            r[key] = info[key]
        r['n_incoming'] = len(info['incoming'])
        for ce in range(3):
            d, c = info['NPRACH_detection'][ce], info['NPRACH_collision'][ce]
            assert len(d) == len(c)
            r['n_occ'][ce] = len(d)
            k = min(len(d), MAX_OCC)
            r['det_hist'][ce][:k] = d[:k]
            r['col_hist'][ce][:k] = c[:k]
        r['nprach_conf'] = [info['NPRACH conf'][k] for k in NPRACH_ITEMS]
        if hist_out is not None:
            hist_out.append((np.asarray(info['incoming'], dtype=np.float64), np.asarray(info['delays'], dtype=np.int32),
                             np.asarray(info['service_times'], dtype=np.int32)))
    if rng is not None:
        r['draws_lo'] = rng.ctr & 0xFFFFFFFF
        r['draws_hi'] = rng.ctr >> 32
    return r


class RefCell:
    """NodeB-level driver of one reference cell: reset() / step(action23) -> record."""

    def __init__(self, conf, seed, n_carriers=1, log_events=False, capture_effect=False):
        m = load_reference(n_carriers)
        self.m = m
        self.rng = PhiloxShim(seed)
        self.node, self.perf, self.population, self.carrier = m['create_system'](self.rng, dict(conf))
        if capture_effect:
            # create_system never passes capture_effect to Channel (system_creator.py:51); the flag is a plain attribute read at every
            # preamble detection (channel.py:126, :158), so it is switched on on the constructed channel
            self.node.m.channel.capture_effect = True
        self.hists = []
        self.events = [] if log_events else None
        if log_events:
            fel = m['em'].fel
            orig = fel.pop_next
            names = ['NPDCCH_arrival', 'NPRACH_update', 'RAR_window_end', 'NPRACH_start', 'NPRACH_end',
                     'Msg3', 'NPUSCH_end', 'MAC_timer', 'Backoff_end']

            def pop_next():
                t, ev = orig()
                self.events.append((t, names.index(ev.type)))
                return t, ev
            fel.pop_next = pop_next
            self._restore = (fel, orig)

    def _rec(self, info, reward):
        if info['state'] == 'RAR_window':
            info = dict(info)
            info['_full_conf'] = self.node.NPRACH_conf
        return info_to_record(info, reward, self.rng, self.hists)

    def reset(self):
        return self._rec(self.node.reset(), 0.0)

    def step(self, action):
        reward, _, info = self.node.step(list(int(a) for a in action))
        return self._rec(info, reward)

    def counters(self):
        p = self.perf
        return {'ue_arrivals': list(p.ue_arrivals), 'ue_departures': list(p.ue_departures), 'backoffs': list(p.backoffs),
                'ue_connections': p.ue_connections, 'error_count': p.error_count, 'attempts_count': p.attempts_count,
                'total_bits': p.total_bits, 'draws': self.rng.ctr, 'ue_count': self.population.ue_count,
                'av_delay': float(p.av_delay)}

    def close(self):
        if self.events is not None:
            fel, orig = self._restore
            fel.pop_next = orig
