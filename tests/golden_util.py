"""Loading and comparing the golden fixtures of tests/golden (made by tools/gen_golden.py from the
unmodified Python reference)."""
import glob
import json
import os

import numpy as np

from nb_iot_environment_b200.record import RECORD_DTYPE, FLOAT_FIELDS, INT_FIELDS, NON_PARITY_FIELDS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))) if not n.startswith(('wrappers_', 'long_', 'mb_agent')))

# float fields are compared to the reference's numpy/libm/scipy arithmetic: nbiot_math.h agrees to ~1e-15,
# the contract (BASELINE.json north_star) is 1e-5 relative
REF_RTOL = 1e-9


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.meta = json.loads(bytes(z['meta']).decode())
        assert self.meta['record_itemsize'] == RECORD_DTYPE.itemsize
        assert self.meta['record_descr'] == str(RECORD_DTYPE.descr), 'record layout changed: regenerate the fixtures'
        raw = z['records']
        self.records = np.ascontiguousarray(raw).view(RECORD_DTYPE).reshape(raw.shape[0], raw.shape[1])
        self.actions = z['actions'].astype(np.int32)
        self.conf = self.meta['conf']
        self.n_carriers = self.meta['n_carriers']
        self.seeds = self.meta['seeds']
        self.decisions = self.meta['decisions']
        self.counters = self.meta['counters']
        self.capture_effect = bool(self.meta.get('capture_effect', False))     # Channel(capture_effect=True): a constructor argument, not a conf key
        self.z = z

    def hist(self, si):
        """per NPRACH_update decision: (incoming, delays, service_times) in order of occurrence"""
        inc, il = self.z[f'incoming_{si}'], self.z[f'incoming_len_{si}']
        dl, sv, ll = self.z[f'delays_{si}'], self.z[f'service_{si}'], self.z[f'delays_len_{si}']
        out, a, b = [], 0, 0
        for n_i, n_d in zip(il, ll):
            out.append((inc[a:a + n_i], dl[b:b + n_d], sv[b:b + n_d]))
            a += n_i
            b += n_d
        return out

    def events(self, si):
        return self.z[f'events_{si}']

    def stats(self, si):
        """PerfMonitor histories of a statistics=True case (delay, connection, access int32; th float64), or None"""
        if f'stat_delay_{si}' not in self.z.files:
            return None
        return {'delay': self.z[f'stat_delay_{si}'], 'connection': self.z[f'stat_conn_{si}'], 'access': self.z[f'stat_access_{si}'], 'th': self.z[f'stat_th_{si}']}


    def pm_lists(self, si):
        """PerfMonitor's sample lists of a traces / statistics case as the reference holds them (lists of tuples / values), or None"""
        pre, suf = 'pm_', f'_{si}'
        keys = [k for k in self.z.files if k.startswith(pre) and k.endswith(suf)]
        if not any(k.startswith('pm_RAR_in') or k.startswith('pm_rar_detections') for k in keys):
            return None
        out = {}
        for k in keys:
            name, v = k[len(pre):-len(suf)], self.z[k]
            vals = [tuple(int(x) for x in row) for row in v] if v.ndim == 2 else [int(x) for x in v]
            if name[-2:] in ('_0', '_1', '_2') and not name.endswith('history'):
                out.setdefault(name[:-2], [None, None, None])[int(name[-1])] = vals
            else:
                out[name] = vals
        return out


def assert_pm_lists(got, want, what=''):
    """got: samples_to_lists(...) of a cell; want: Golden.pm_lists(si). Every list the golden holds must be equal."""
    for k, v in want.items():
        assert got[k] == v, f'{what} PerfMonitor.{k} differs'


def record_diff(a, b, rtol=0.0, skip=NON_PARITY_FIELDS):
    """list of (field, a, b) that differ: integers exactly, floats exactly (rtol=0) or within rtol"""
    out = []
    for n in INT_FIELDS:
        if n in skip:
            continue
        if not np.array_equal(a[n], b[n]):
            out.append((n, a[n], b[n]))
    for n in FLOAT_FIELDS:
        if rtol == 0.0:
            same = np.array_equal(a[n], b[n])
        else:
            same = np.allclose(a[n], b[n], rtol=rtol, atol=1e-12)
        if not same:
            out.append((n, a[n], b[n]))
    return out
