"""ctypes binding of the CPU oracle (oracle/nbiot_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; nothing under nb_iot_environment_b200/ imports this module."""
import ctypes
import os
import subprocess

import numpy as np

from nb_iot_environment_b200.record import RECORD_DTYPE, TRACE_DTYPE, NbiotConf, make_conf

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libnbiot_oracle.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(NbiotConf), ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_stats.restype = ctypes.c_int
        L.orc_get_stats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_get_step_list.restype = ctypes.c_int
        L.orc_get_step_list.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_get_samples.restype = ctypes.c_int
        L.orc_get_samples.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_get_hist.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 2
        L.orc_get_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_set_event_log.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_event_log_len.argtypes = [ctypes.c_void_p]
        L.orc_run_fixed.restype = ctypes.c_longlong
        L.orc_run_fixed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_rng_u01.restype = ctypes.c_double
        L.orc_rng_u01.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.orc_rng_normal.restype = ctypes.c_double
        L.orc_rng_normal.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
        L.orc_rng_bounded.restype = ctypes.c_uint64
        L.orc_rng_bounded.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64]
        L.orc_rng_pair.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
        L.orc_bench.restype = ctypes.c_double
        L.orc_bench.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_bench_records.restype = ctypes.c_double
        L.orc_bench_records.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        assert L.orc_record_size() == RECORD_DTYPE.itemsize
        assert L.orc_conf_size() == ctypes.sizeof(NbiotConf)
        _LIB = L
    return _LIB


_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        d = os.path.join(_HERE, '..', 'nb_iot_environment_b200', 'data')
        _TABLES = (np.fromfile(os.path.join(d, 'lut2_u16.bin'), dtype='<u2'),
                   np.fromfile(os.path.join(d, 'mcs_u8.bin'), dtype='u1'))
        assert _TABLES[0].size == 7 * 8 * 501 and _TABLES[1].size == 500 * 30 * 3
    return _TABLES


class OracleCell:
    """One simulated cell: NodeB-level reset()/step(action23) like reference system/node_b.py:163-245."""

    def __init__(self, conf, seed, n_carriers=1, capture_effect=False):
        self.L = lib()
        self.cconf = conf if isinstance(conf, NbiotConf) else make_conf(conf, n_carriers=n_carriers, capture_effect=capture_effect)
        lut, mcs = tables()
        self.h = self.L.orc_create(ctypes.byref(self.cconf), ctypes.c_uint64(seed), lut.ctypes.data, mcs.ctypes.data)
        self.rec = np.zeros(1, dtype=RECORD_DTYPE)
        self._evlog = None

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.orc_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.orc_reset(self.h, self.rec.ctypes.data)
        return self.rec[0].copy()

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.int32)
        assert a.size == 23
        self.L.orc_step(self.h, a.ctypes.data, self.rec.ctypes.data)
        return self.rec[0].copy()

    def run_fixed(self, action, until_time):
        a = np.ascontiguousarray(action, dtype=np.int32)
        n = self.L.orc_run_fixed(self.h, a.ctypes.data, int(until_time), self.rec.ctypes.data)
        return int(n), self.rec[0].copy()

    def hist(self, cap=4096):
        inc = np.zeros(cap, dtype=np.float64); d = np.zeros(cap, dtype=np.int32); s = np.zeros(cap, dtype=np.int32)
        ni = ctypes.c_int32(); nd = ctypes.c_int32()
        self.L.orc_get_hist(self.h, inc.ctypes.data, d.ctypes.data, s.ctypes.data, cap, ctypes.byref(ni), ctypes.byref(nd))
        return inc[:ni.value].copy(), d[:nd.value].copy(), s[:nd.value].copy()

    def stats(self, cap=1 << 16):
        """statistics histories since reset (perf_monitor.py:61-84): delay, connection, access (int32) and th = bits / delay"""
        out = {}
        buf = np.zeros(cap, dtype=np.int32)
        for which, name in enumerate(('delay', 'connection', 'access', 'bits')):
            n = self.L.orc_get_stats(self.h, which, buf.ctypes.data, cap)
            assert n <= cap
            out[name] = buf[:n].copy()
        out['th'] = out['bits'].astype(np.float64) / out['delay'].astype(np.float64)      # acked_data / delay (perf_monitor.py:169)
        return out

    def step_lists(self, cap=4096):
        """the per-epoch info lists in full: (receptions dict, errors, unfit, access dict) as the reference's info holds them"""
        out = []
        buf = np.zeros(cap, dtype=np.int32)
        for which in range(6):
            n = self.L.orc_get_step_list(self.h, which, buf.ctypes.data, cap)
            out.append(buf[:n].tolist())
        return dict(zip(out[0], out[1])), out[2], out[3], dict(zip(out[4], out[5]))

    def samples(self, cap=1 << 18):
        """PerfMonitor's sample lists since reset, in append order (TRACE_DTYPE)"""
        buf = np.zeros(cap, dtype=TRACE_DTYPE)
        n = self.L.orc_get_samples(self.h, buf.ctypes.data, cap)
        assert n <= cap
        return buf[:n].copy()

    def counters(self):
        out = np.zeros(16, dtype=np.int64); av = ctypes.c_double()
        self.L.orc_get_counters(self.h, out.ctypes.data, ctypes.byref(av))
        return {'ue_arrivals': out[0:3].tolist(), 'ue_departures': out[3:6].tolist(), 'backoffs': out[6:9].tolist(),
                'ue_connections': int(out[9]), 'error_count': int(out[10]), 'attempts_count': int(out[11]),
                'total_bits': int(out[12]), 'draws': int(out[13]), 'pops': int(out[14]), 'ue_count': int(out[15]),
                'av_delay': av.value}

    def enable_event_log(self, cap=1 << 20):
        self._evlog = np.zeros((cap, 2), dtype=np.int32)
        self.L.orc_set_event_log(self.h, self._evlog.ctypes.data, cap)

    def event_log(self):
        return self._evlog[:self.L.orc_event_log_len(self.h)].copy()


def bench(conf, ctrl, seeds, actions=None, steps=0, until_time=-1, n_threads=1):
    """Time n cells on n_threads host threads (oracle/nbiot_oracle.c: orc_bench). actions: uint8[steps][n][ext_k]
    external-agent actions, or None to run on the internal agents until `until_time`.
    Returns (seconds of the slowest thread inside the step loop, simulated subframes, NodeB steps)."""
    L = lib()
    lut, mcs = tables()
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    sf, ns = ctypes.c_longlong(), ctypes.c_longlong()
    ap = None
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        assert actions.shape == (steps, len(seeds), ctrl.ext_k)
        ap = actions.ctypes.data
    sec = L.orc_bench(ctypes.byref(conf), ctypes.byref(ctrl), lut.ctypes.data, mcs.ctypes.data, seeds.ctypes.data, len(seeds), ap,
                      int(steps), int(until_time), int(n_threads), ctypes.byref(sf), ctypes.byref(ns))
    return sec, sf.value, ns.value


def run_batch(conf, ctrl, seeds, actions=None, steps=0, until_time=-1, n_threads=None):
    """Simulate a batch on host threads and return the last decision record of every cell (RECORD_DTYPE[n])."""
    import os
    L = lib()
    lut, mcs = tables()
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    recs = np.zeros(len(seeds), dtype=RECORD_DTYPE)
    sf, ns = ctypes.c_longlong(), ctypes.c_longlong()
    ap = None
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        assert actions.shape == (steps, len(seeds), ctrl.ext_k)
        ap = actions.ctypes.data
    L.orc_bench_records(ctypes.byref(conf), ctypes.byref(ctrl), lut.ctypes.data, mcs.ctypes.data, seeds.ctypes.data, len(seeds), ap,
                        int(steps), int(until_time), int(n_threads or os.cpu_count() or 1), ctypes.byref(sf), ctypes.byref(ns), recs.ctypes.data)
    return recs
