"""Test-side restatement of the reference's Controller (controller.py:23-288) with fixed-action internal
agents (control_agents.py:83-113), driving ONE CPU-oracle cell. Used to check the device-resident
controller tables of the CUDA path."""
import numpy as np

from nb_iot_environment_b200 import parameters as par
from nb_iot_environment_b200.controller import ControllerTables
from nb_iot_environment_b200.record import STATE_NAMES, record_to_info
from oracle.oracle_py import OracleCell


class Stalled(Exception):
    pass


class OracleController:
    def __init__(self, conf, seed, agents_conf=None, ext_agent=-1, n_carriers=1, capture_effect=False):
        self.cell = OracleCell(conf, seed, n_carriers=n_carriers, capture_effect=capture_effect)
        self.t = ControllerTables(agents_conf, ext_agent, n_carriers)
        self.action = np.zeros(par.N_actions, dtype=np.int32)
        for ag in self.t.agents:                       # _initialize_action (controller.py:63-69)
            self.action[ag.a_mask] = ag.fixed_action
        self.rec = None
        self.hist = None

    @property
    def ext(self):
        return self.t.ext_agent

    def _update_action(self, aid):
        ag = self.t.agents[aid]
        self.action[ag.a_mask] = ag.fixed_action

    def _node_step(self):
        self.rec = self.cell.step(self.action)
        if self.rec['state'] == 1:
            self.hist = self.cell.hist()

    def get_obs(self, items=None):                     # controller.py:105-120
        items = self.ext.obs_items if items is None else items
        info = record_to_info(self.rec)
        obs = []
        for item in items:
            n_, m_ = par.obs_dict[item]
            if n_ == 1:
                obs.append(min(1.0, info.get(item, 0) / m_))
            else:
                for v in info.get(item, [0] * n_):
                    obs.append(min(1.0, v / m_))
        return np.asarray(obs, dtype=float)

    def reset(self):
        self.rec = self.cell.reset()
        self.hist = self.cell.hist()
        return self.get_obs() if self.ext is not None else None

    def step(self, a):                                 # controller.py:234-288
        ext = self.ext
        self.action[ext.a_mask] = a
        nxt = ext.next
        while nxt > 0:
            self._update_action(nxt)
            nxt = self.t.agents[nxt].next
        self._node_step()
        state = STATE_NAMES[self.rec['state']]
        steps = 1
        while state not in ext.states:
            if steps >= 65536:                             # NBIOT_STALL_STEPS: the reference loops on for ever, the product faults
                raise Stalled()
            for aid in self.t.agents_ids[state]:
                self._update_action(aid)
            self._node_step()
            steps += 1
            state = STATE_NAMES[self.rec['state']]
        for aid in self.t.agents_ids[state]:
            if aid == ext.id:
                break
            self._update_action(aid)
        return self.get_obs(), float(self.rec['reward'])

    def run_until(self, time):                         # controller.py:185-195 with single_step :156-165
        n = 0
        while self.rec['time'] < time:
            state = STATE_NAMES[self.rec['state']]
            for aid in self.t.agents_ids[state]:
                self._update_action(aid)
            self._node_step()
            n += 1
        return n
