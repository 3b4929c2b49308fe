"""Host side of the multi-agent split: turns the reference's agent dictionaries
(control_agents.py:29-71) into the write tables the kernels apply on device
(struct nbiot_ctrl, csrc/nbiot_state.h), following reference controller.py:

  * every agent owns the indices `a_mask` of the shared 23-int action (control_agents.py:86-92);
  * Controller._initialize_action writes every agent's fixed action once, in agent order (:63-69);
  * Controller.step scatters the external agent's action, then runs its `next` chain (:234-247);
  * Controller.to_next_step lets the agents of each visited state act until a state of the external
    agent is reached, where only the agents listed before it act (:262-288).

Internal agents are fixed-action agents (the reference's DummyAgent); their actions can be changed
with `set_action` exactly like DummyAgent.set_action (:100-103)."""
import ctypes

import numpy as np

from . import parameters as par
from .record import STATE_IDS


class NbiotCtrl(ctypes.Structure):
    _fields_ = [('ext_state_mask', ctypes.c_uint32), ('ext_k', ctypes.c_int32),
                ('ext_a_mask', ctypes.c_uint8 * 24), ('ext_post', ctypes.c_int8 * 24),
                ('state_writes', (ctypes.c_int8 * 24) * 4), ('init_action', ctypes.c_uint8 * 24),
                ('full_writes', (ctypes.c_int8 * 24) * 4)]


class AgentSpec:
    """attrs of the reference's agent protocol: id, action_items, obs_items, next, states -> a_mask, a_max, fixed_action"""

    def __init__(self, conf, n_carriers=1):
        self.id = conf['id']
        self.action_items = list(conf['action_items'])
        self.obs_items = list(conf['obs_items'])
        self.next = conf['next']
        self.states = list(conf['states'])
        self.a_mask = [par.control_items[n] for n in self.action_items]
        self.a_max = [par.control_max_values[n][n_carriers - 1] for n in self.action_items]
        self.fixed_action = [par.control_default_values[n] for n in self.action_items]

    def set_action(self, values):
        for i, name in enumerate(self.action_items):
            if name in values:
                self.fixed_action[i] = int(values[name])


class ControllerTables:
    def __init__(self, agents_conf=None, ext_agent=-1, n_carriers=1):
        agents_conf = par.agents_conf if agents_conf is None else agents_conf
        self.agents = [AgentSpec(c, n_carriers) for c in agents_conf]
        for i, a in enumerate(self.agents):
            if a.id != i:
                raise ValueError('agents must be listed in id order (reference controller.py indexes self.agents[id])')
        self.n_carriers = n_carriers
        self.agents_ids = {}                                   # controller.py:43-54
        for a in self.agents:
            for s in a.states:
                self.agents_ids.setdefault(s, []).append(a.id)
        self.set_ext_agent(ext_agent)

    def set_ext_agent(self, ext_agent):
        self.ext_agent = self.agents[ext_agent] if ext_agent > -1 else None

    @property
    def obs_dim(self):
        if self.ext_agent is None:
            return 0
        return sum(par.obs_dict[i][0] for i in self.ext_agent.obs_items)

    def initial_action(self):
        a = [0] * 24
        for ag in self.agents:                                  # Controller._initialize_action
            for idx, v in zip(ag.a_mask, ag.fixed_action):
                a[idx] = v
        return a

    def build(self, all_states_external=False):
        """-> NbiotCtrl.  all_states_external: every NodeB decision returns to the host with the full
        23-vector as the action (NodeB.step-level driving, used by the parity tests)."""
        c = NbiotCtrl()
        for s in range(4):
            for i in range(24):
                c.state_writes[s][i] = -1
                c.full_writes[s][i] = -1
        # Controller.single_step / run_until (controller.py:156-195): every agent of the state acts, whoever is registered as external
        for sname, ids in self.agents_ids.items():
            for aid in ids:
                ag = self.agents[aid]
                for idx, v in zip(ag.a_mask, ag.fixed_action):
                    c.full_writes[STATE_IDS[sname]][idx] = v
        for i in range(24):
            c.ext_post[i] = -1
        init = self.initial_action()
        for i in range(24):
            c.init_action[i] = init[i]
        if all_states_external:
            c.ext_state_mask = 0xF
            c.ext_k = par.N_actions
            for i in range(par.N_actions):
                c.ext_a_mask[i] = i
            return c
        ext = self.ext_agent
        c.ext_state_mask = 0
        c.ext_k = 0
        if ext is not None:
            for s in ext.states:
                c.ext_state_mask |= 1 << STATE_IDS[s]
            c.ext_k = len(ext.a_mask)
            for i, idx in enumerate(ext.a_mask):
                c.ext_a_mask[i] = idx
            nxt = ext.next
            while nxt > 0:                                      # controller.py:244-247 (`> 0`, not `>= 0`)
                ag = self.agents[nxt]
                for idx, v in zip(ag.a_mask, ag.fixed_action):
                    c.ext_post[idx] = v
                nxt = ag.next
        for sname, ids in self.agents_ids.items():
            s = STATE_IDS[sname]
            for aid in ids:
                if ext is not None and sname in ext.states and aid == ext.id:
                    break                                       # controller.py:283-288: agents before the external one
                ag = self.agents[aid]
                for idx, v in zip(ag.a_mask, ag.fixed_action):
                    c.state_writes[s][idx] = v
        return c

    # ---- spaces (controller.py:77-103) as plain data; gymnasium is optional at run time
    def action_nvec(self):
        return np.asarray(self.ext_agent.a_max, dtype=np.int64) + 1

    def obs_layout(self):
        """[(item name, length, norm)] of the external agent's observation (controller.py:105-120)."""
        return [(i, par.obs_dict[i][0], par.obs_dict[i][1]) for i in self.ext_agent.obs_items]
