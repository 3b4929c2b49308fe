"""Test double with the BatchedNBIoTEnv interface (num_envs, device, action_nvec, reset, step(actions, active=),
record_field) built from N CPU-oracle cells behind the test-side Controller restatement. It lets the batched
wrappers be checked against the reference's wrappers without a GPU."""
import numpy as np
import torch

from nb_iot_environment_b200.record import RECORD_DTYPE
from oracle_controller import OracleController


class _Infos:
    def __init__(self, recs):
        self.records = recs


class OracleVectorEnv:
    def __init__(self, conf, agents_conf, ext_agent, seeds, n_carriers=1):
        self.ctrls = [OracleController(conf, s, agents_conf=agents_conf, ext_agent=ext_agent, n_carriers=n_carriers) for s in seeds]
        self.num_envs = len(seeds)
        self.device = torch.device('cpu')
        self.action_nvec = self.ctrls[0].t.action_nvec()
        self.obs_dim = self.ctrls[0].t.obs_dim
        self._obs = np.zeros((self.num_envs, self.obs_dim))
        self._rew = np.zeros(self.num_envs)
        self._recs = np.zeros(self.num_envs, dtype=RECORD_DTYPE)
        self._false = torch.zeros(self.num_envs, dtype=torch.bool)

    def reset(self):
        for i, c in enumerate(self.ctrls):
            self._obs[i] = c.reset()
            self._recs[i] = c.rec
        return torch.from_numpy(self._obs.copy()), _Infos(self._recs.copy())

    def step(self, actions, active=None):
        a = np.asarray(actions.cpu() if isinstance(actions, torch.Tensor) else actions).reshape(self.num_envs, -1)
        act = np.ones(self.num_envs, dtype=bool) if active is None else np.asarray(active.cpu() if isinstance(active, torch.Tensor) else active, dtype=bool)
        for i, c in enumerate(self.ctrls):
            if act[i]:
                self._obs[i], self._rew[i] = c.step(a[i])
                self._recs[i] = c.rec
        return torch.from_numpy(self._obs.copy()), torch.from_numpy(self._rew.copy()), self._false, self._false, _Infos(self._recs.copy())

    def record_field(self, name):
        return torch.from_numpy(np.ascontiguousarray(self._recs[name]))
