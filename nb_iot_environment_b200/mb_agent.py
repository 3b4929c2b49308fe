"""The reference's model-based NPRACH agent for a batch of cells, on the device.

`BatchedNPRACHAgent` mirrors agent_nprach.NPRACH_THAgent (reference agent_nprach.py:41-146): same constructor
arguments (`buffer_size`, `margin`, `no_th`), `set_parameter(margin)`, `get_action(...)` returning the 8 action items
of agent 3 (th_C1, th_C0, sc_C0..2, period_C0..2) -- for every cell of a BatchedNBIoTEnv at once, as CUDA tensors. It
reads the NPRACH_update decision records where they are (HBM) through the C ABI of include/nbiot_agent.h; the model
tables come from nb_iot_environment_b200/data/mb_agent.npz (tools/make_agent_tables.py). No CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from . import _cabi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'mb_agent.npz')
METRICS = ('departures', 'NPRACH_occupation', 'service_times', 'beta')      # evaluate_MBRL.py:44


def load_tables():
    z = np.load(_DATA)
    t = {k: z[k] for k in z.files}
    if t['ml_est'].shape != (4, 128, 64) or t['rate_conf'].shape != (186, 8) or t['cfg_rate'].shape[0] != 3:
        raise RuntimeError('mb_agent.npz is corrupt: rerun tools/make_agent_tables.py')
    return t


class BatchedNPRACHAgent:
    def __init__(self, env, metrics=METRICS, buffer_size=100, margin=1.8, no_th=False):
        self.env, self.L = env, _cabi.load()
        self.N = env.num_envs
        self.metrics = list(metrics)
        self.samples = {m: [] for m in self.metrics}         # per step: a tensor [N] on the device (reference: a float per step)
        t = load_tables()
        K = 32
        cfg_rate = np.full((3, K), np.inf); cfg_conf = np.zeros((3, K, 2), np.uint8)
        k0 = t['cfg_rate'].shape[1]
        cfg_rate[:, :k0] = t['cfg_rate']; cfg_conf[:, :k0] = t['cfg_conf']
        self._keep = [np.ascontiguousarray(t['ml_est'], dtype=np.int16), np.ascontiguousarray(cfg_rate), np.ascontiguousarray(cfg_conf),
                      np.ascontiguousarray(t['rate_conf'], dtype=np.uint8)]
        T = _cabi.MbTables()
        T.ml_est, T.cfg_rate, T.cfg_conf, T.rate_conf = (a.ctypes.data for a in self._keep)
        T.cfg_n[:] = [int(x) for x in t['cfg_n']]
        T.m3_det[:] = [float(x) for x in t['m3_det']]
        T.m3_ce[:] = [int(x) for x in t['m3_ce']]
        T.min_value, T.max_value, T.step = (float(x) for x in t['consts'])
        T.period_list[:] = [int(x) for x in t['period_list']]
        T.th_default[:] = [int(x) for x in t['th_default']]
        with torch.cuda.device(env.device):
            self.state = torch.zeros(self.N * _cabi.MB_STATE_BYTES, dtype=torch.uint8, device=env.device)     # caller-owned: a checkpoint of the agents
            self._actions = torch.zeros((self.N, 8), dtype=torch.uint8, device=env.device)
            self._samples = torch.zeros((self.N, 4), dtype=torch.float64, device=env.device)
            self._margin = torch.full((self.N,), float(margin), dtype=torch.float64, device=env.device)
            h = ctypes.c_void_p()
            _cabi.check(self.L.nbiot_mb_agent_create(env._h, ctypes.byref(T), int(buffer_size), int(bool(no_th)), float(margin),
                                                     self.state.data_ptr(), env._stream(), ctypes.byref(h)))
        self._h = h

    # ---- reference API
    def set_parameter(self, margin):
        """margin: a float, or one value per cell (tensor / array)"""
        if isinstance(margin, torch.Tensor):
            self._margin.copy_(margin.to(device=self.env.device, dtype=torch.float64).reshape(self.N))
        elif np.ndim(margin) == 0:
            self._margin.fill_(float(margin))
        else:
            self._margin.copy_(torch.as_tensor(np.asarray(margin, dtype=np.float64).reshape(self.N)))

    def get_action(self, obs=None, r=None, info=None, action=None):
        """uint8[N, 8] on the device: th_C1, th_C0, sc_C0, sc_C1, sc_C2, period_C0, period_C1, period_C2
        (the arguments are accepted for signature compatibility; the agent reads the decision records itself)"""
        with torch.cuda.device(self.env.device):
            _cabi.check(self.L.nbiot_mb_agent_act(self._h, self._margin.data_ptr(), self._actions.data_ptr(), self._samples.data_ptr(), self.env._stream()))
            for j, m in enumerate(METRICS):
                if m in self.samples:
                    self.samples[m].append(self._samples[:, j].clone())
        return self._actions

    # ---- agent memory
    def _field(self, dtype, offset, count):
        esz = torch.empty((), dtype=dtype).element_size()
        flat = self.state.view(dtype)
        return torch.as_strided(flat, (self.N, count), (_cabi.MB_STATE_BYTES // esz, 1), offset // esz)

    @property
    def msg3_detection(self):
        return self._field(torch.float64, 0, 3)

    @property
    def lambda_estimate(self):
        return self._field(torch.float64, 24, 1)[:, 0]

    @property
    def k(self):
        return self._field(torch.int32, 44, 1)[:, 0]

    @property
    def conf(self):
        return self._field(torch.uint8, 48, 6)

    def faults(self):
        c = ctypes.c_longlong(0)
        _cabi.check(self.L.nbiot_mb_agent_faults(self._h, ctypes.byref(c)))
        return int(c.value)

    def close(self):
        if self._h:
            self.L.nbiot_mb_agent_destroy(self._h)
            self._h = None
