"""Batched restatements of the reference's gym wrappers (reference wrappers.py), SURVEY.md section 8(f) rank 1.

Each wrapper holds per-instance state as tensors on the environment's device and follows the reference wrapper
instance by instance:

  NPRACHWrapper          wrappers.py:119-208   threshold-pair action space (105 feasible pairs), validity check of the
                                               NPRACH packing (carrier.compute_offsets), invalid -> replay the previous
                                               action with reward -1, reward / norm otherwise
  BasicSchedulerWrapper  wrappers.py:47-116    impossible UE index -> 10 * min_reward WITHOUT stepping the cell,
                                               penalty when the chosen MCS did not fit
  DiscreteActions        wrappers.py:355-363   flat index -> MultiDiscrete tuple
  RARWrapper             wrappers.py:366-421   msg3 / departure based reward shaping

They work on any object with the BatchedNBIoTEnv interface (num_envs, device, action_nvec, reset, step(actions,
active=), record_field)."""
import itertools

import numpy as np
import torch

from . import parameters as par

# ---- parameters.py:340-367
N_sc_list = [3, 6, 9, 12]                                  # parameters.py:340 (the second assignment wins)
period_list = [80, 160, 240, 320, 640, 1280]               # parameters.py:339 (max_p = 5)
NPRACH_periodicity_list = [80, 160, 240, 320, 640, 1280, 2560]
NPRACH_N_sf_list = [6, 12, 23, 45, 90, 180, 359, 717]      # ceil(5.6 * N_rep), parameters.py:343
thr_to_n_rep = {-116: 3, -115: 3, -114: 3, -113: 2, -112: 2, -111: 1, -110: 1}
no_rep_th = -109


def compute_NPRACH_sf(th_C1, th_C0):
    """parameters.compute_NPRACH_sf (parameters.py:345-367)"""
    min_th = par.threshold_list[0]
    sf_c2 = NPRACH_N_sf_list[thr_to_n_rep[min_th]]
    rep_C1 = thr_to_n_rep[max(min_th, th_C1)] if th_C1 < no_rep_th else 0
    rep_C0 = thr_to_n_rep[max(min_th, th_C0)] if th_C0 < no_rep_th else 0
    sf_c1, sf_c0 = NPRACH_N_sf_list[rep_C1], NPRACH_N_sf_list[rep_C0]
    if th_C0 == 0 or th_C1 == th_C0:
        sf_c0 = 0
    return [sf_c0, sf_c1, sf_c2]


def offsets_valid(periods, N_sfs, N_scs):
    """validity flag of carrier.compute_offsets (carrier.py:115-252)"""
    min_length = sum(N_sfs)
    critical = [i for i, p in enumerate(periods) if p == 240]
    normal = [i for i, p in enumerate(periods) if p != 240]
    if len(critical) == 1 and max(N_sfs) > 80:
        i, j, k = critical[0], normal[0], normal[1]
        if N_scs[i] + max(N_scs[j], N_scs[k]) > 12:
            return False
        min_length = max(N_sfs[j] + N_sfs[k], N_sfs[i]) if sum(N_scs) > 12 else N_sfs[2]
    elif len(critical) == 2 and max(N_sfs) > 80:
        i, j, k = critical[0], critical[1], normal[0]
        if N_scs[k] + max(N_scs[i], N_scs[j]) > 12:
            return False
        min_length = max(N_sfs[i] + N_sfs[j], N_sfs[k]) if sum(N_scs) > 12 else N_sfs[2]
    elif sum(N_scs) > 12:
        if N_scs[2] + max(N_scs[1], N_scs[0]) <= 12:
            min_length = N_sfs[2] + N_sfs[0]
        elif N_scs[2] + N_scs[1] <= 12:
            min_length = N_sfs[2] + N_sfs[0]
        elif N_scs[1] + N_scs[0] <= 12:
            min_length = N_sfs[2] + N_sfs[1]
        elif N_scs[2] + N_scs[0] <= 12:
            min_length = N_sfs[2] + N_sfs[1]
    else:
        min_length = max(N_sfs[2], N_sfs[1] + N_sfs[0])
    min_periodicity = min(p for p in NPRACH_periodicity_list if p >= min_length)
    return min(periods) >= min_periodicity


_VALID_TABLE = None


def nprach_validity_table():
    """bool[105, 4, 4, 4, 6, 6, 6]: validity of every wrapper action (threshold pair, sc_C0..2, period_C0..2)"""
    global _VALID_TABLE
    if _VALID_TABLE is None:
        t = np.zeros((len(par.th_pairs), 4, 4, 4, 6, 6, 6), dtype=bool)
        sfs = [compute_NPRACH_sf(a, b) for a, b in par.th_pairs]
        for (s0, s1, s2) in itertools.product(range(4), repeat=3):
            nsc = [N_sc_list[s0], N_sc_list[s1], N_sc_list[s2]]
            for (p0, p1, p2) in itertools.product(range(6), repeat=3):
                periods = [period_list[p0], period_list[p1], period_list[p2]]
                for ti, nsf in enumerate(sfs):
                    t[ti, s0, s1, s2, p0, p1, p2] = offsets_valid(periods, nsf, nsc)
        _VALID_TABLE = t
    return _VALID_TABLE


def to_discrete(dimensions):
    """wrappers.to_discrete (wrappers.py:26-28): row i = the i-th tuple of itertools.product(range(d) for d in dimensions)"""
    return np.asarray(list(itertools.product(*[range(int(d)) for d in dimensions])), dtype=np.int64)


class _Wrapper:
    def __init__(self, env):
        self.env = env
        self.num_envs = env.num_envs
        self.device = env.device

    def __getattr__(self, name):
        if name in ('env',):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, seed=None, options=None):
        return self.env.reset()

    def _div(self, x, d):
        """x / d as IEEE division (a CUDA tensor divided by a Python scalar is multiplied by the rounded reciprocal)"""
        return x / torch.full_like(x, float(d))

    def _actions(self, a, k):
        return torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(self.device).long().reshape(self.num_envs, k)


class NPRACHWrapper(_Wrapper):
    """reference NPRACH_wrapper (wrappers.py:119-208) for the external agent of evaluate_RL.py:70-77
    (action items th_C1, th_C0, sc_C0, sc_C1, sc_C2, period_C0, period_C1, period_C2)."""

    def __init__(self, env, discrete=False, use_default_action=False, norm=100):
        super().__init__(env)
        self.default_action = use_default_action
        self.norm = norm
        self.n = 0
        self.action = torch.tensor([5, 8, 2, 2, 2, 2, 2, 2], device=self.device).repeat(self.num_envs, 1)    # wrappers.py:131
        nvec = np.asarray(env.action_nvec, dtype=np.int64)[1:].copy()
        nvec[0] = len(par.th_pairs)
        self.action_nvec = nvec
        self.discrete = discrete
        if discrete:
            self.actions = torch.from_numpy(to_discrete(nvec)).to(self.device)
        self._valid = torch.from_numpy(nprach_validity_table()).to(self.device)
        self._pair_idx = torch.tensor([par.th_indexes[p] for p in par.th_pairs], device=self.device)           # [105, 2]

    def step(self, action):
        self.n += 1
        a = self._actions(action, 1 if self.discrete else 7)
        if self.discrete:
            a = self.actions[a[:, 0]]
        valid = self._valid[a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4], a[:, 5], a[:, 6]]
        transformed = torch.cat([self._pair_idx[a[:, 0]], a[:, 1:]], dim=1)
        applied = torch.where(valid[:, None], transformed, self.action)
        obs, reward, terminated, truncated, infos = self.env.step(applied)
        reward = torch.where(valid, self._div(reward, self.norm), torch.full_like(reward, -1.0))
        if not self.default_action:
            self.action = applied
        self.last_valid = valid
        return obs, reward, terminated, truncated, infos


class NPRACHWrapperTraces(NPRACHWrapper):
    """reference NPRACH_wrapper_traces (wrappers.py:210-255), the wrapper scripts/nprach/evaluate_RL.py:141 runs: after every step one
    sample per metric and cell is kept -- info[metric], or the mean of it when it is a list (0 for an empty one) -- on the device.
    `samples[metric]` is a list with one float64[N] tensor per step; `samples_array(metric)` stacks them to [steps, N], the per-cell
    columns being what the reference's evaluator collects per run (evaluate_RL.py:172-177). The log file / flag file of the reference
    class are host conveniences and are not restated."""

    _PERIOD_LISTS = {'incoming': (0, 'n_incoming'), 'delays': (1, 'departures'), 'service_times': (2, 'departures')}

    def __init__(self, env, metrics, discrete=False, use_default_action=False, norm=100):
        super().__init__(env, discrete=discrete, use_default_action=use_default_action, norm=norm)
        self.samples = {m: [] for m in metrics}

    def _period_mean(self, which, count_field):
        base = self.env
        off, cap = base._hist_off[which], base._hist_off[3]
        N = self.num_envs
        if which == 0:
            v = base.state[off:off + N * cap * 8].view(torch.float64).view(N, cap)
        else:
            v = base.state[off:off + N * cap * 4].view(torch.int32).view(N, cap).double()
        n = base.record_field(count_field).long().clamp(max=cap)
        keep = torch.arange(cap, device=self.device)[None, :] < n[:, None]
        total = torch.where(keep, v, torch.zeros_like(v)).sum(dim=1)
        return torch.where(n > 0, total / n.clamp(min=1).double(), torch.zeros_like(total))

    def _sample(self, metric):
        if metric in self._PERIOD_LISTS:
            return self._period_mean(*self._PERIOD_LISTS[metric])
        v = self.env.record_field(metric).double()
        return v.mean(dim=1) if v.dim() == 2 else v            # a fixed-length list of the info (ratios per CE level ...): its mean

    def step(self, action):
        out = super().step(action)
        for metric, lst in self.samples.items():
            lst.append(self._sample(metric).clone())
        return out

    def samples_array(self, metric):
        return torch.stack(self.samples[metric]) if self.samples[metric] else torch.zeros((0, self.num_envs), dtype=torch.float64, device=self.device)


class BasicSchedulerWrapper(_Wrapper):
    """reference BasicSchedulerWrapper (wrappers.py:84-116); with select_only=True the plain BasicWrapper (:47-82)."""

    def __init__(self, env, select_only=False):
        super().__init__(env)
        self.select_only = select_only
        self.n = 0
        self.min_reward = torch.full((self.num_envs,), -100.0, dtype=torch.float64, device=self.device)
        self.n_users = torch.ones(self.num_envs, dtype=torch.int64, device=self.device)       # info = {'ues': [0]} before the first step
        self.n_unfit = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self.action_nvec = env.action_nvec

    def step(self, action):
        self.n += 1
        k = len(self.action_nvec)
        a = self._actions(action, k)
        impossible = (self.n_users > 1) & (a[:, 0] >= self.n_users)
        stepped = ~impossible
        obs, env_reward, terminated, truncated, infos = self.env.step(a, active=stepped)
        env_reward = env_reward.clone()
        self.min_reward = torch.where(stepped & (env_reward < self.min_reward), env_reward, self.min_reward)
        reward = torch.where(impossible, 10.0 * self.min_reward, env_reward)
        # self.info is only refreshed for the instances that stepped
        self.n_users = torch.where(stepped, self.env.record_field('n_sel').long(), self.n_users)
        if not self.select_only:
            self.n_unfit = torch.where(stepped, self.env.record_field('n_unfit').long(), self.n_unfit)
            reward = torch.where(self.n_unfit > 0, reward + self.min_reward, reward)
        self.last_impossible = impossible
        return obs, reward, terminated, truncated, infos


class BasicWrapper(BasicSchedulerWrapper):
    """reference BasicWrapper (wrappers.py:47-82): the agent only selects the UE; no penalty for grants that do not fit"""

    def __init__(self, env):
        super().__init__(env, select_only=True)


class DiscreteActions(_Wrapper):
    """reference DiscreteActions (wrappers.py:355-363)"""

    def __init__(self, env):
        super().__init__(env)
        self.actions = torch.from_numpy(to_discrete(env.action_nvec)).to(self.device)
        self.n_actions = int(self.actions.shape[0])

    def step(self, act):
        a = self._actions(act, 1)
        return self.env.step(self.actions[a[:, 0]])


class RARWrapper(_Wrapper):
    """reference RAR_wrapper (wrappers.py:366-421): reward from msg3 detections (or departures) and failed msg3"""

    def __init__(self, env, reward_criteria='msg3'):
        super().__init__(env)
        self.reward_criteria = reward_criteria
        z = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device)
        self.total_msg_t_1, self.total_departures_1 = z.clone(), z.clone()
        self.n_steps = 0

    def step(self, action):
        obs, _, terminated, truncated, infos = self.env.step(action)
        f = self.env.record_field
        total_msg_t = f('RAR_detected').long().sum(dim=1)
        detections = (total_msg_t - self.total_msg_t_1).clamp(min=0)
        self.total_msg_t_1 = total_msg_t
        total_failed = f('RAR_failed').long().sum(dim=1)
        total_departures = f('total_departures').long()
        departures = (total_departures - self.total_departures_1).clamp(min=0)
        self.total_departures_1 = total_departures
        gain = departures if self.reward_criteria == 'departures' else detections
        reward = torch.clamp(torch.clamp(gain.double() / 4, max=1.0) - total_failed.double() / 4, min=-1.0)
        reward = torch.where(f('valid_CE_control') == 0, torch.full_like(reward, -1.0), reward)
        self.n_steps += 1
        return obs, reward, terminated, truncated, infos


class NPRACHAgentWrapper(_Wrapper):
    """reference NPRACH_agent_wrapper (wrappers.py:257-318): the learner only picks the margin (beta) of the model-based
    configurator; the agent (mb_agent.BatchedNPRACHAgent) turns it into the 8 NPRACH action items on the device.
    Observation = the selected items of the environment's observation followed by the last `obs_window` arrival-rate
    estimates of the agent; reward = environment reward / norm."""

    def __init__(self, env, agent, bounds=(1.4, 3.4), reduced_observation=True, discrete=True, n_actions=30, obs_items=(11, 12), obs_window=1, norm=100):
        super().__init__(env)
        self.agent = agent
        self.samples = agent.samples
        self.discrete = discrete
        self.bounds = bounds
        if discrete:
            self.action_values = torch.from_numpy(np.linspace(bounds[0], bounds[1], n_actions)).to(self.device)   # wrappers.py:266
        self.reduced_observation = reduced_observation
        self.obs_items = torch.as_tensor(list(obs_items), device=self.device, dtype=torch.long)
        self.obs_window = obs_window
        self.arrival_history = torch.zeros((self.num_envs, obs_window), dtype=torch.float64, device=self.device)   # deque([0] * obs_window)
        self.norm = norm
        self.n = 0

    def _reduce(self, obs):
        if not self.reduced_observation:
            return obs
        return torch.cat([obs[:, self.obs_items], self.arrival_history], dim=1)

    def reset(self, seed=None, options=None):
        obs, infos = self.env.reset()
        self.obs = self._reduce(obs)
        return self.obs, infos

    def step(self, action):
        self.n += 1
        if self.discrete:
            margin = self.action_values[self._actions(action, 1)[:, 0]]
        else:
            margin = torch.as_tensor(action, dtype=torch.float64, device=self.device).reshape(self.num_envs, -1)[:, 0]
        self.agent.set_parameter(margin)
        conf = self.agent.get_action()
        obs, reward, terminated, truncated, infos = self.env.step(conf)
        self.conf = conf
        self.r = self._div(reward, self.norm)
        if self.reduced_observation:
            self.arrival_history = torch.cat([self.arrival_history[:, 1:], self.agent.lambda_estimate[:, None].clone()], dim=1)
        self.obs = self._reduce(obs)
        return self.obs, self.r, terminated, truncated, infos


class NPRACHAgentWrapperTraces(NPRACHAgentWrapper):
    """reference NPRACH_agent_wrapper_traces (wrappers.py:321-353): the running mean of every metric the agent samples, updated
    incrementally after each step as the reference does (avg += (sample - avg) / n), one float64[N] tensor per metric on the
    device. The log file of the reference class is a host convenience and is not restated."""

    def __init__(self, env, agent, **kw):
        super().__init__(env, agent, **kw)
        self.averages = {m: torch.zeros(self.num_envs, dtype=torch.float64, device=self.device) for m in self.samples}

    def step(self, action):
        out = super().step(action)
        for metric, samples in self.samples.items():
            if samples:
                avg = self.averages[metric]
                self.averages[metric] = avg + self._div(samples[-1] - avg, self.n)
        return out
