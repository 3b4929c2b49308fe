"""ctypes driver of tests/_hostemu (the CPU build of the device simulation core). Debug aid for tests only."""
import ctypes
import os
import subprocess

import numpy as np

from nb_iot_environment_b200.controller import ControllerTables, NbiotCtrl
from nb_iot_environment_b200.record import RECORD_DTYPE, TRACE_DTYPE, NbiotConf, make_conf
from oracle.oracle_py import tables

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_hostemu')
_LIBS = {}


def lib(lanes):
    if lanes not in _LIBS:
        subprocess.check_call(['make', '-s', '-C', _HERE])
        L = ctypes.CDLL(os.path.join(_HERE, '_build', f'libemu{lanes}.so'))      # lanes: 1, 32, '1t' (1 lane, task chain) or '32h' (32 lanes, event pool filled from the top)
        L.emu_create.restype = ctypes.c_void_p
        L.emu_create.argtypes = [ctypes.POINTER(NbiotConf), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(NbiotCtrl)]
        L.emu_destroy.argtypes = [ctypes.c_void_p]
        L.emu_set_ctrl.argtypes = [ctypes.c_void_p, ctypes.POINTER(NbiotCtrl)]
        L.emu_set_evlog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.emu_reset.argtypes = [ctypes.c_void_p]
        L.emu_advance.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.emu_records.restype = ctypes.c_void_p
        L.emu_stats.restype = ctypes.c_int
        L.emu_stats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.emu_records.argtypes = [ctypes.c_void_p]
        L.emu_trace.restype = ctypes.c_int
        L.emu_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.emu_counters.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        assert L.emu_lanes() == {'1t': 1, '1tg': 1, '1b': 1, '32h': 32}.get(lanes, lanes) and L.emu_tasks() == (1 if lanes in ('1t', '1tg') else 0)
        _LIBS[lanes] = L
    return _LIBS[lanes]


class EmuBatch:
    def __init__(self, conf, seeds, n_carriers=1, lanes=1, ctrl=None, capture_effect=False, **caps):
        self.L = lib(lanes)
        self.cconf = make_conf(conf, n_carriers=n_carriers, capture_effect=capture_effect, **caps)
        self.n = len(seeds)
        self.seeds = np.asarray(seeds, dtype=np.uint64)
        lut, mcs = tables()
        self.tabs = ControllerTables(n_carriers=n_carriers)
        self.ctrl = ctrl if ctrl is not None else self.tabs.build(all_states_external=True)
        self.h = self.L.emu_create(ctypes.byref(self.cconf), self.n, self.seeds.ctypes.data, lut.ctypes.data, mcs.ctypes.data, ctypes.byref(self.ctrl))
        self.evlog = None

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.emu_destroy(self.h)
            self.h = None

    def records(self):
        p = self.L.emu_records(self.h)
        buf = (ctypes.c_uint8 * (self.n * RECORD_DTYPE.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=RECORD_DTYPE).copy()

    def enable_evlog(self, cap=1 << 18):
        self.evlog = np.zeros((cap, 2), dtype=np.int32)
        self.L.emu_set_evlog(self.h, self.evlog.ctypes.data, cap)

    def reset(self):
        self.L.emu_reset(self.h)
        return self.records()

    def advance(self, ext_actions=None, until_time=-1, max_steps=1 << 30):
        steps = np.zeros(self.n, dtype=np.int32)
        if ext_actions is not None:
            a = np.ascontiguousarray(ext_actions, dtype=np.uint8)
            assert a.shape == (self.n, self.ctrl.ext_k)
            self.L.emu_advance(self.h, a.ctypes.data, until_time, max_steps, steps.ctypes.data)
        else:
            self.L.emu_advance(self.h, None, until_time, max_steps, steps.ctypes.data)
        return self.records(), steps

    def stats(self, i, cap=1 << 16):
        out = {}
        buf = np.zeros(cap, dtype=np.int32)
        for which, name in enumerate(('delay', 'connection', 'access', 'bits')):
            n = self.L.emu_stats(self.h, i, which, buf.ctypes.data, cap)
            out[name] = buf[:min(n, cap, self.cconf.stat_cap)].copy()
            out['n'] = n
        return out

    def samples(self, i, cap=1 << 16):
        buf = np.zeros(cap, dtype=TRACE_DTYPE)
        n = self.L.emu_trace(self.h, i, buf.ctypes.data, cap)
        return buf[:min(n, cap, self.cconf.trace_cap)].copy(), n

    def counters(self, i):
        out = np.zeros(16, dtype=np.int64); av = ctypes.c_double()
        self.L.emu_counters(self.h, i, out.ctypes.data, ctypes.byref(av))
        return {'ue_arrivals': out[0:3].tolist(), 'ue_departures': out[3:6].tolist(), 'backoffs': out[6:9].tolist(),
                'ue_connections': int(out[9]), 'error_count': int(out[10]), 'attempts_count': int(out[11]),
                'total_bits': int(out[12]), 'draws': int(out[13]), 'events': int(out[14]), 'ue_count': int(out[15]),
                'av_delay': av.value}
