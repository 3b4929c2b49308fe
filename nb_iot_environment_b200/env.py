"""BatchedNBIoTEnv: N independent NB-IoT cells behind the reference's Gymnasium-style interface.

One instance behaves like `gym.make('gym_system:System-v1', system=Controller(node, agents))` of the
reference (gym-system/gym_system/environment.py:22-37 -> controller.py:132-143, :234-288 ->
system/node_b.py:163-245): `reset()` returns the observation of the first decision (always an NPRACH
update at t = 0), `step(actions)` scatters the external agent's action items into the shared 23-int
action, lets the internal fixed-action agents drive the cell until the external agent's state is
reached again and returns `(obs, reward, terminated=False, truncated=False, infos)`.

All N cells advance inside CUDA kernels (csrc/nbiot_kernels.cu) through the C ABI of include/nbiot.h.
PyTorch only provides device memory and streams. Observations and rewards stay on the device as
float64 tensors; `infos` is a lazy view that copies the decision records to the host on first use."""
import ctypes

import numpy as np
import torch

from . import _cabi
from . import parameters as par
from .controller import ControllerTables
from .record import RECORD_DTYPE, TRACE_DTYPE, STATE_IDS, FAULT_FATAL_MASK, make_conf, record_to_info, samples_to_lists

_ALL = 0xF
_SCHED = (1 << STATE_IDS['Scheduling']) | (1 << STATE_IDS['RAR_window_end'])
_RAR = 1 << STATE_IDS['RAR_window']
_UPD = 1 << STATE_IDS['NPRACH_update']
_CONF_POS = {n: i for i, n in enumerate(par.NPRACH_items)}

# obs item -> (record field, element offset in the field, kind, NodeB states whose info holds the key)
_OBS_SOURCES = {
    'total_ues': ('total_ues', 0, 0, _ALL), 'carrier_state': ('carrier_busy', 0, 2, _ALL),
    'connection_time': ('connection_time', 0, 0, _SCHED), 'loss': ('loss', 0, 1, _SCHED), 'sinr': ('sinr', 0, 1, _SCHED),
    'buffer': ('buffer', 0, 0, _SCHED),
    'detection_ratios': ('detection_ratios', 0, 1, _UPD), 'colision_ratios': ('colision_ratios', 0, 1, _UPD),
    'msg3_detection': ('msg3_detection', 0, 1, _UPD), 'NPUSCH_occupation': ('NPUSCH_occupation', 0, 1, _UPD),
    'NPRACH_occupation': ('NPRACH_occupation', 0, 1, _UPD), 'av_delay': ('av_delay', 0, 1, _UPD),
    'distribution': ('distribution', 0, 1, _UPD),
    'ues_per_CE': ('ues_per_CE', 0, 0, _RAR), 'RAR_in': ('RAR_in', 0, 0, _RAR), 'RAR_sent': ('RAR_sent', 0, 0, _RAR),
    'RAR_detected': ('RAR_detected', 0, 0, _RAR), 'RAR_ids': ('RAR_ids', 0, 0, _RAR), 'NPDCCH_sf_left': ('NPDCCH_sf_left', 0, 0, _RAR),
}
for _n in ('sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2', 'th_C0', 'th_C1'):
    _OBS_SOURCES[_n] = ('nprach_conf', _CONF_POS[_n], 0, _RAR)


def build_obs_items(obs_items):
    """obs item names -> ctypes array of nbiot_obs_item_t (Controller.get_obs, controller.py:105-120)."""
    arr = (_cabi.ObsItem * max(1, len(obs_items)))()
    for k, name in enumerate(obs_items):
        field, elem, kind, mask = _OBS_SOURCES[name]
        length, norm = par.obs_dict[name]
        base = RECORD_DTYPE.fields[field][1]
        esize = {0: 4, 1: 8, 2: 1}[kind]
        arr[k].offset = base + elem * esize
        arr[k].kind = kind
        arr[k].length = length
        arr[k].state_mask = mask
        arr[k].norm = float(norm)
    return arr


class LazyInfos:
    """Sequence of per-instance info dicts (reference keys), materialised on demand."""

    def __init__(self, env):
        self._env = env
        self._records = None
        self._hist = None
        self._spill = None
        self._epoch = env._epoch          # the step this view belongs to: the records live in the state buffer and move on with it

    def _check_fresh(self):
        if self._epoch != self._env._epoch:
            raise _cabi.NbiotError('this infos object belongs to an earlier step and was not read before the environment moved on: the reference '
                                   'returns a snapshot dict, here the records are read on first use -- read infos (or infos.records) before the next step / reset')

    @property
    def records(self):
        if self._records is None:
            self._check_fresh()
            self._records = self._env.records()
        return self._records

    def _hists(self):
        if self._hist is None:
            self._check_fresh()
            self._hist = self._env.period_lists()
        return self._hist

    def __len__(self):
        return self._env.num_envs

    def _spills(self):
        if self._spill is None:
            self._check_fresh()
            self._spill = self._env.step_list_spill()
        return self._spill

    def __getitem__(self, i):
        rec = self.records[i]
        hist = None
        spill = None
        if max(int(rec['n_rx']), int(rec['n_err']), int(rec['n_unfit']), int(rec['n_acc'])) > 16:
            spill = self._spills()[i]              # a per-epoch list longer than the record's 16 entries: the rest is in the side buffer
        if rec['state'] == STATE_IDS['NPRACH_update']:
            inc, dl, sv = self._hists()
            cap = inc.shape[1]
            ni, nd = min(int(rec['n_incoming']), cap), min(int(rec['departures']), cap)
            hist = (inc[i, :ni].tolist(), dl[i, :nd].tolist(), sv[i, :nd].tolist())
        return record_to_info(rec, hist, spill)

    def field(self, name):
        return self.records[name]


class BatchedNBIoTEnv:
    """conf: the reference's conf dict (system/system_creator.py:24-44); agents_conf: list of agent dicts
    (control_agents.py:29-71; None = the reference's default six agents); ext_agent: index of the agent
    that is driven from outside (-1: none, the cells run on their internal agents). kernel: which kernel runs the long advances -- None / 'auto' (by batch size), 'warp', 'thread', 'thread_voted'
    (results are bit-identical; include/nbiot_conf.h says which load favours which). capture_effect: Channel(capture_effect=True) of the
    reference (channel.py:124-126; create_system never sets it, so it is a constructor argument here too)."""

    def __init__(self, conf, agents_conf=None, ext_agent=-1, num_envs=1, seeds=None, n_carriers=1, device=0,
                 ue_cap=None, grid_w=1024, hist_cap=256, all_states_external=False, stat_cap=None, trace_cap=None, capture_effect=False, kernel=None):
        if not torch.cuda.is_available():
            raise _cabi.NbiotError('BatchedNBIoTEnv needs a CUDA device: there is no CPU fallback')
        self.L = _cabi.load()
        self.num_envs = int(num_envs)
        self.device = torch.device('cuda', device)
        self.conf = dict(conf)
        self.n_carriers = n_carriers
        self.cconf = make_conf(conf, n_carriers=n_carriers, ue_cap=ue_cap, grid_w=grid_w, hist_cap=hist_cap, stat_cap=stat_cap, trace_cap=trace_cap, capture_effect=capture_effect, kernel=kernel)
        self.tables = ControllerTables(agents_conf, ext_agent, n_carriers)
        self.all_states_external = bool(all_states_external)
        self.ctrl = self.tables.build(all_states_external=self.all_states_external)
        seeds = np.arange(self.num_envs, dtype=np.uint64) if seeds is None else np.asarray(seeds, dtype=np.uint64)
        if seeds.shape != (self.num_envs,):
            raise ValueError('seeds must have one entry per instance')
        self.seeds = seeds
        nbytes = self.L.nbiot_state_bytes(ctypes.byref(self.cconf), self.num_envs)
        # the caller-owned simulation state: saving this tensor (plus conf) is a checkpoint
        self.state = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        from .tables import load_tables
        lut, mcs = load_tables()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _cabi.check(self.L.nbiot_create(ctypes.byref(self.cconf), self.num_envs, self.device.index, self.state.data_ptr(), nbytes,
                                            seeds.ctypes.data, lut.ctypes.data, mcs.ctypes.data, ctypes.byref(self.ctrl), st,
                                            ctypes.byref(self._h)))
        self._epoch = 0                   # advances whenever the cells move (step / reset / run_until / load_state): stamps LazyInfos
        self._set_obs_items()
        self._obs = torch.zeros((self.num_envs, max(1, self.obs_dim)), dtype=torch.float64, device=self.device)
        self._reward = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self._stats = torch.zeros(_cabi.N_STATS, dtype=torch.int64, device=self.device)
        self._false = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        rp = ctypes.c_void_p()
        _cabi.check(self.L.nbiot_records_ptr(self._h, ctypes.byref(rp)))
        self._rec_off = rp.value - self.state.data_ptr()
        ip, dp, sp, cap = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        _cabi.check(self.L.nbiot_hist_ptrs(self._h, ctypes.byref(ip), ctypes.byref(dp), ctypes.byref(sp), ctypes.byref(cap)))
        base = self.state.data_ptr()
        self._hist_off = (ip.value - base, dp.value - base, sp.value - base, cap.value)
        stp, scap = ctypes.c_void_p(), ctypes.c_int()
        _cabi.check(self.L.nbiot_stat_ptrs(self._h, ctypes.byref(stp), ctypes.byref(scap)))
        self.stat_cap = scap.value
        self._stat_off = (stp.value - base) if stp.value else None
        spp, lcap, scap2 = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        _cabi.check(self.L.nbiot_spill_ptr(self._h, ctypes.byref(spp), ctypes.byref(lcap), ctypes.byref(scap2)))
        self._spill_off, self._spill_cap = spp.value - base, scap2.value
        trp, tcap = ctypes.c_void_p(), ctypes.c_int()
        _cabi.check(self.L.nbiot_trace_ptrs(self._h, ctypes.byref(trp), ctypes.byref(tcap)))
        self.trace_cap = tcap.value
        self._trace_off = (trp.value - base) if trp.value else None

    # ------------------------------------------------------------------ spaces (controller.py:77-103)
    def _set_obs_items(self):
        ext = self.tables.ext_agent
        names = list(ext.obs_items) if (ext is not None and not self.all_states_external) else []
        self.obs_items = names
        self._items = build_obs_items(names)
        self._n_items = len(names)
        self.obs_dim = sum(par.obs_dict[n][0] for n in names)

    @property
    def action_nvec(self):
        """MultiDiscrete(a_max + 1) of the external agent (Discrete when it has a single item)."""
        if self.all_states_external:
            # one entry per INDEX of the shared vector (two pairs of items share an index, parameters.py:159-185): the largest maximum + 1
            mx = [0] * par.N_actions
            for name, idx in par.control_items.items():
                mx[idx] = max(mx[idx], par.control_max_values[name][self.n_carriers - 1])
            return np.asarray(mx, dtype=np.int64) + 1
        return self.tables.action_nvec()

    @property
    def single_action_space(self):
        import gymnasium
        nvec = self.action_nvec
        return gymnasium.spaces.MultiDiscrete(nvec) if len(nvec) > 1 else gymnasium.spaces.Discrete(int(nvec[0]))

    @property
    def single_observation_space(self):
        import gymnasium
        return gymnasium.spaces.Box(low=0.0, high=1.0, shape=(self.obs_dim,), dtype=float)

    def set_ext_agent(self, ext_agent, obs_items=None):
        """Controller.set_ext_agent (controller.py:122-130)."""
        self.tables.set_ext_agent(ext_agent)
        self.ctrl = self.tables.build(all_states_external=self.all_states_external)
        _cabi.check(self.L.nbiot_set_controller(self._h, ctypes.byref(self.ctrl), self._stream()))
        self._set_obs_items()
        self._obs = torch.zeros((self.num_envs, max(1, self.obs_dim)), dtype=torch.float64, device=self.device)

    def set_internal_action(self, agent_id, values):
        """DummyAgent.set_action (control_agents.py:100-103) for an internal agent."""
        self.tables.agents[agent_id].set_action(values)
        self.ctrl = self.tables.build(all_states_external=self.all_states_external)
        _cabi.check(self.L.nbiot_set_controller(self._h, ctypes.byref(self.ctrl), self._stream()))

    # ------------------------------------------------------------------ gym interface
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _gather(self):
        """Controller.get_obs for the batch. The observation and reward tensors are the environment's own buffers, rewritten by
        the next step (a learner that keeps them across steps clones them -- documented aliasing, as torch vector envs do);
        pass fresh=True to step() for new tensors every call."""
        _cabi.check(self.L.nbiot_gather_obs(self._h, self._items, self._n_items, self._obs.data_ptr(), self._reward.data_ptr(), self._stream()))
        return self._obs[:, :self.obs_dim]

    def _actions_u8(self, actions, k):
        """int[N, k] -> uint8 on the device, refusing values a uint8 cast would wrap into another (possibly valid) action"""
        if isinstance(actions, torch.Tensor):
            a = actions.to(device=self.device).reshape(self.num_envs, k)
            if a.dtype != torch.uint8:
                if a.numel() and (int(a.min()) < 0 or int(a.max()) > 255):
                    raise ValueError('action values must be in 0..255 (and below action_nvec)')
                a = a.to(torch.uint8)
            return a.contiguous()
        a = np.asarray(actions).reshape(self.num_envs, k)
        if a.dtype != np.uint8:
            if a.size and (a.min() < 0 or a.max() > 255):
                raise ValueError('action values must be in 0..255 (and below action_nvec)')
            a = a.astype(np.uint8)
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def reset(self, seed=None, options=None):
        """(obs[N, D], infos). `seed` is ignored, as in the reference (environment.py:22-28)."""
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_reset(self._h, self._stream()))
            self._epoch += 1
            obs = self._gather()
        return obs, LazyInfos(self)

    def step(self, actions, active=None, fresh=False, validate=False):
        """actions: int[N, k] (k = number of action items of the external agent) as a CUDA tensor (stays on
        the device) or a host array (copied). Returns (obs, reward, terminated, truncated, infos).
        active: optional bool[N]; instances with False are not stepped at all (their observation, reward and
        record stay as they were) -- used by wrappers that answer without calling the environment."""
        k = self.ctrl.ext_k
        with torch.cuda.device(self.device):
            a = self._actions_u8(actions, k)
            if validate:                    # one compare + any() on the device, and a host sync: off by default, the kernel faults on what the reference raises on
                nvec = torch.as_tensor(np.asarray(self.action_nvec), device=self.device)
                if bool((a.to(torch.int64) >= nvec).any()):
                    raise ValueError('action outside action_nvec')
            if active is None:
                _cabi.check(self.L.nbiot_advance(self._h, a.data_ptr(), -1, 1 << 30, self._stream()))
            else:
                m = torch.as_tensor(active).to(device=self.device, dtype=torch.uint8).reshape(self.num_envs).contiguous()
                _cabi.check(self.L.nbiot_advance_masked(self._h, a.data_ptr(), m.data_ptr(), -1, 1 << 30, self._stream()))
                self._last_mask = m
            self._last_actions = a          # keep alive until the kernel has consumed it
            self._epoch += 1
            obs = self._gather()
            rew = self._reward
            if fresh:
                obs, rew = obs.clone(), rew.clone()
        return obs, rew, self._false, self._false, LazyInfos(self)

    def step_host(self, actions_host, obs_out=None, reward_out=None):
        """The same step through host buffers in ONE C-ABI call (nbiot_step_host): actions are copied
        host->device, the observation and reward device->host, and the call returns synchronised."""
        k = self.ctrl.ext_k
        a = np.asarray(actions_host).reshape(self.num_envs, k)
        if a.dtype != np.uint8:
            if a.size and (a.min() < 0 or a.max() > 255):
                raise ValueError('action values must be in 0..255 (and below action_nvec)')
            a = a.astype(np.uint8)
        a = np.ascontiguousarray(a)
        if obs_out is None:
            obs_out = np.empty((self.num_envs, max(1, self.obs_dim)), dtype=np.float64)
        if reward_out is None:
            reward_out = np.empty(self.num_envs, dtype=np.float64)
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_step_host(self._h, a.ctypes.data, -1, 1 << 30, self._items, self._n_items,
                                               obs_out.ctypes.data, reward_out.ctypes.data, self._stream()))
            self._epoch += 1
        return obs_out[:, :self.obs_dim], reward_out

    def run_until(self, time, max_steps=1 << 30):
        """Controller.run_until (controller.py:185-195): every cell runs on its internal agents until its
        clock reaches `time`."""
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_advance(self._h, None, int(time), int(max_steps), self._stream()))
            self._epoch += 1

    def node_step(self, actions23):
        """NodeB.step-level driving (all_states_external=True): one decision per call with the full 23-vector."""
        assert self.all_states_external
        with torch.cuda.device(self.device):
            a = self._actions_u8(actions23, par.N_actions)
            _cabi.check(self.L.nbiot_advance(self._h, a.data_ptr(), -1, 1, self._stream()))
            self._last_actions = a
            self._epoch += 1

    # ------------------------------------------------------------------ results
    def record_field(self, name):
        """one field of the decision records as a torch tensor ON THE DEVICE (a strided view of the state buffer;
        int32 / float64 fields only) -- lets wrappers read info values without a host round trip."""
        dt, off = RECORD_DTYPE.fields[name][0], RECORD_DTYPE.fields[name][1]
        base, shape = dt.base, dt.shape
        tdt = {np.dtype('<i4'): torch.int32, np.dtype('<f8'): torch.float64}[base]
        esz = base.itemsize
        n_el = int(np.prod(shape)) if shape else 1
        flat = self.state[self._rec_off:self._rec_off + self.num_envs * RECORD_DTYPE.itemsize].view(tdt)
        stride = RECORD_DTYPE.itemsize // esz
        v = torch.as_strided(flat, (self.num_envs, n_el), (stride, 1), flat.storage_offset() + off // esz)
        return v if shape else v[:, 0]

    def records(self):
        """decision records of all instances as a numpy structured array (device -> host copy)."""
        n = self.num_envs * RECORD_DTYPE.itemsize
        raw = self.state[self._rec_off:self._rec_off + n].cpu().numpy()
        return raw.view(RECORD_DTYPE).copy()

    # ------------------------------------------------------------------ statistics histories (conf['statistics'])
    def stat_counts(self):
        """int32[N] on the device: departures recorded in the histories since reset"""
        out = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_stat_counts(self._h, out.data_ptr(), self._stream()))
        return out

    def samples(self):
        """the sample ring of every cell in order of occurrence: a list of (TRACE_DTYPE array, samples since reset); a cell with more
        samples than trace_cap keeps the last trace_cap of them"""
        if not self.trace_cap:
            raise _cabi.NbiotError("the batch was created without sample lists: pass conf['traces']=True and / or conf['statistics']=True")
        N, cap = self.num_envs, self.trace_cap
        cnt = torch.zeros(N, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_trace_counts(self._h, cnt.data_ptr(), self._stream()))
        cnt = cnt.cpu().numpy()
        raw = self.state[self._trace_off:self._trace_off + N * cap * TRACE_DTYPE.itemsize].cpu().numpy().view(TRACE_DTYPE).reshape(N, cap)
        return [(raw[i, :n].copy() if n <= cap else np.roll(raw[i], -(n % cap)).copy(), int(n)) for i, n in enumerate(cnt)]

    def traces(self):
        """PerfMonitor's `traces` lists (perf_monitor.py:61-69, :111-140) of every cell under the reference's attribute names:
        rar_detections / rar_errors / rar_collisions / nprach_arrivals / nprach_detections per CE level and the t_* lists over all levels,
        each a list of (value, t). Needs conf['traces']."""
        if not self.cconf.trace_mask & 3:
            raise _cabi.NbiotError("the batch was created without traces: pass conf['traces']=True")
        out = []
        for rec, n in self.samples():
            d = samples_to_lists(rec)
            d = {k: d[k] for k in ('rar_detections', 'rar_errors', 'rar_collisions', 't_rar_detections', 'nprach_arrivals', 'nprach_detections',
                                   't_nprach_arrivals', 't_nprach_detections')}
            d['truncated'] = n > self.trace_cap
            out.append(d)
        return out

    def histories(self):
        """PerfMonitor.delay_history / connection_history / access_history / th_history (perf_monitor.py:61-84, :161-171) of every cell:
        a list of dicts of numpy arrays, in departure order (what experiments_RL.py:139-140 saves as `delay` and `connection`), plus the
        other `statistics` lists: RAR_in / RAR_attmp / RAR_sent / RAR_detected and npusch_delivered per CE level, msg3 / NPUSCH /
        NPRACH_resources_history (:71-79, :125-129, :164-168, :185-202).
        A cell with more departures than stat_cap keeps the last stat_cap of them ('truncated': True)."""
        if not self.stat_cap:
            raise _cabi.NbiotError("the batch was created without statistics: pass conf['statistics']=True or stat_cap")
        N, cap = self.num_envs, self.stat_cap
        cnt = self.stat_counts().cpu().numpy()
        raw = self.state[self._stat_off:self._stat_off + N * 4 * cap * 4].cpu().numpy().view(np.int32).reshape(N, 4, cap)
        out = []
        for i in range(N):
            n = int(cnt[i])
            if n <= cap:
                v = raw[i, :, :n]
            else:                                    # ring: the oldest kept entry sits at n % cap
                v = np.roll(raw[i], -(n % cap), axis=1)
            out.append({'delay': v[0].copy(), 'connection': v[1].copy(), 'access': v[2].copy(),
                        'th': v[3].astype(np.float64) / v[0].astype(np.float64), 'served_users': n, 'truncated': n > cap})
        if self.trace_cap and self.cconf.trace_mask & 0x7C:
            for h, (rec, n) in zip(out, self.samples()):
                d = samples_to_lists(rec)
                h.update({k: d[k] for k in ('RAR_in', 'RAR_attmp', 'RAR_sent', 'RAR_detected', 'npusch_delivered', 'msg3_resources_history',
                                            'NPUSCH_resources_history', 'NPRACH_resources_history')})
                h['samples_truncated'] = n > self.trace_cap
        return out

    def save_results(self, path, i=0):
        """np.savez with the reference's keys (experiments_RL.py:139-140, experiments_MAMBRL.py:121-122) for cell i"""
        h = self.histories()[i]
        np.savez(path, delay=h['delay'], connection=h['connection'], served_users=h['served_users'])

    def stat_histogram(self, which='delay', bin_width=100, n_bins=64):
        """int64[n_bins] on the device: histogram of one history over the whole batch (last bin open-ended)"""
        idx = {'delay': 0, 'connection': 1, 'access': 2, 'bits': 3}[which]
        out = torch.zeros(n_bins, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_stat_histogram(self._h, idx, int(bin_width), int(n_bins), out.data_ptr(), self._stream()))
        return out

    def step_list_spill(self):
        """int32[N, 6, max(48, hist_cap)] (host): entries 16.. of the per-epoch info lists (rx id / delay, errors, unfit, access id / delay) of every cell"""
        N, cap = self.num_envs, self._spill_cap
        return self.state[self._spill_off:self._spill_off + N * 6 * cap * 4].cpu().numpy().view(np.int32).reshape(N, 6, cap)

    def period_lists(self):
        """(incoming float64[N, cap], delays int32[N, cap], service_times int32[N, cap]) of the update period
        reported by the last NPRACH_update decision record (node_b.py:403-407)."""
        oi, od, os_, cap = self._hist_off
        N = self.num_envs
        inc = self.state[oi:oi + N * cap * 8].cpu().numpy().view(np.float64).reshape(N, cap)
        dl = self.state[od:od + N * cap * 4].cpu().numpy().view(np.int32).reshape(N, cap)
        sv = self.state[os_:os_ + N * cap * 4].cpu().numpy().view(np.int32).reshape(N, cap)
        return inc, dl, sv

    # ------------------------------------------------------------------ checkpoint / resume
    def save_state(self):
        """Snapshot of the complete simulation state (every instance: clock, event slots, lists, UE records, random
        counters). The reference has no checkpointing; here it is one tensor copy."""
        torch.cuda.synchronize(self.device)
        return self.state.clone()

    def load_state(self, snapshot):
        """Restore a snapshot taken from an environment created with the same conf, capacities and num_envs."""
        if snapshot.numel() != self.state.numel():
            raise ValueError('snapshot size does not match this environment (conf / capacities / num_envs differ)')
        self.state.copy_(snapshot.to(self.device))
        self._epoch += 1
        with torch.cuda.device(self.device):
            self._gather()                  # the cached observation / reward follow the restored records
        torch.cuda.synchronize(self.device)

    def stats_tensor(self):
        """int64[N_STATS] on the device: episode statistics summed over this handle's instances."""
        with torch.cuda.device(self.device):
            _cabi.check(self.L.nbiot_reduce_stats(self._h, self._stats.data_ptr(), self._stream()))
        return self._stats

    def stats(self):
        v = self.stats_tensor().cpu().numpy()
        out = {}
        for idx, name in _cabi.STAT_NAMES.items():
            out[name] = v[idx:idx + 3].tolist() if name in ('ue_arrivals', 'ue_departures', 'backoffs') else int(v[idx])
        return out

    def faults(self):
        return self.records()['fault']

    @property
    def launch_count(self):
        return int(self.L.nbiot_launch_count(self._h))

    def launch_geometry(self):
        b, w, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _cabi.check(self.L.nbiot_launch_geometry(self._h, ctypes.byref(b), ctypes.byref(w), ctypes.byref(s)))
        return {'blocks': b.value, 'warps_per_block': w.value, 'smem_bytes': s.value}

    def kernel_name(self):
        """which simulation kernel advances this batch (picked by batch size; NBIOT_KERNEL overrides)"""
        return self.L.nbiot_kernel_name(self._h).decode() if hasattr(self.L, 'nbiot_kernel_name') else 'warp-per-cell'

    def enable_event_log(self, cap=1 << 16):
        self._evlog = torch.zeros((cap, 2), dtype=torch.int32, device=self.device)
        _cabi.check(self.L.nbiot_set_event_log(self._h, self._evlog.data_ptr(), cap))

    def event_log(self):
        return self._evlog.cpu().numpy()

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            self.L.nbiot_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
