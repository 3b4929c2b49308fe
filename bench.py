#!/usr/bin/env python3
"""bench.py -- throughput of the batched NB-IoT cell simulator.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1]): scripts/nprach of the reference -- M=1000 devices per cell, ratio 0.5
('mixed_05', evaluate_RL.py:30-34,108-114), agents of evaluate_RL.py:46-77 with the NPRACH agent (id 3) external,
4096 cell instances per GPU. A STEP is one external decision of every instance = 3000 simulated subframes per
instance (the NPRACH update period). External actions are uniform-random feasible NPRACH configurations
(threshold pair out of the 105 of parameters.th_pairs, per-CE subcarrier and period indices), a function of
(seed, step, global cell index) alone, so the workload is the same for every --steps / --warmup / --gpus.

One JSON line (rank 0): `value` = simulated subframe-instances/s with inputs resident in HBM, timed with CUDA
events on the launching stream, L2 flushed between timed steps; `e2e` = the same through the host-buffer C-ABI
call (actions host->device, observation + reward device->host every step); `roofline` = algorithmic bytes of
SURVEY.md 8(d) (constants in include/nbiot_roofline.h, L-bar and rho_ev measured on the device) over the sim
kernel's CUDA-event time, against the measured HBM peak; `issue` = the ceiling the kernel is actually under
(instruction issue); `cpu_baseline` = the UNMODIFIED reference environment as a multiprocess vector env on this
box's host cores (oracle/ref_vector_env.py over the staged oracle/_ref/), `cpu_port` = the C oracle on all host
threads. The same line carries the other BASELINE.json configurations, each measured in this run with its own
parity guards (no faulted instance, no truncated list): `config5` (the sweep configuration the 1e10 target is
quoted on), `config2` (configs[2], NPUSCH scenario, 65 536 cells), `config3` (configs[3], two carriers, bursty).
`--impl reference` prints the CPU arm alone. A run with a faulted instance exits non-zero.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 4096
SF_PER_STEP = 3000
CONF = {'ratio': 0.5, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
AGENTS = [
    {'id': 0, 'action_items': ['id', 'Imcs', 'Nrep', 'carrier', 'delay', 'sc'], 'obs_items': [], 'next': -1, 'states': ['Scheduling']},
    {'id': 1, 'action_items': ['ce_level', 'rar_Imcs', 'Nrep'], 'obs_items': [], 'next': -1, 'states': ['RAR_window']},
    {'id': 2, 'action_items': ['rar_window', 'mac_timer', 'transmax', 'panchor', 'backoff'], 'obs_items': [], 'next': -1, 'states': ['RAR_window_end']},
    {'id': 3, 'action_items': ['th_C1', 'th_C0', 'sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2'],
     'obs_items': ['detection_ratios', 'colision_ratios', 'msg3_detection', 'NPRACH_occupation', 'av_delay', 'distribution'],
     'next': -1, 'states': ['NPRACH_update']},
]
EXT_AGENT = 3
CAPS = dict(ue_cap=CONF['M'], grid_w=int(os.environ.get('NBIOT_BENCH_GRID_W', '2048')), hist_cap=512)   # live UEs <= M (closed population, ue_generator.py:38-42)
METRIC = 'simulated subframe-instances/sec'
UNIT = 'subframe-instances/s'
WORKLOAD = ('configs[1]: NPRACH control scenario (scripts/nprach, M=1000, ratio=0.5, NPRACH agent external, '
            'uniform-random feasible NPRACH configurations), 4096 instances per GPU, 1 step = 3000 subframes per instance')


def ext_actions(steps, n, seed, first_cell=0):
    """uint8[steps][n][8]: th_C1, th_C0, sc_C0..2, period_C0..2 (action items of agent 3). The action of cell g at step s is a
    function of (seed, s, g) alone -- one generator per step, cells drawn in global order -- so the workload does not change with
    --steps, --warmup, the number of ranks or the batch size."""
    from nb_iot_environment_b200 import parameters as par
    pairs = np.asarray([par.th_indexes[p] for p in par.th_pairs], dtype=np.uint8)       # 105 feasible (th_C1 <= th_C0) pairs
    a = np.zeros((steps, n, 8), dtype=np.uint8)
    for s in range(steps):
        rng = np.random.Generator(np.random.Philox(key=[seed, s]))
        r = rng.integers(0, 1 << 32, size=(first_cell + n, 8), dtype=np.uint64)[first_cell:]
        a[s, :, 0:2] = pairs[(r[:, 0] * len(pairs)) >> 32]
        a[s, :, 2:5] = (r[:, 2:5] * 4) >> 32
        a[s, :, 5:8] = (r[:, 5:8] * 6) >> 32
    return a


def roofline_constants():
    text = open(os.path.join(ROOT, 'include', 'nbiot_roofline.h')).read()
    return {k: int(v) for k, v in re.findall(r'#define NBIOT_RL_(\w+) (\d+)', text)}


def b_alg(lbar, rho, carriers=1):
    c = roofline_constants()
    return 2 * c['H_HOT'] + 2 * c['B_COL'] * carriers + lbar * c['U_HOT'] + rho * (2 * c['H_EVT'] + 2 * c['K_UE'] * c['U_FULL'])


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def _run_nvml(self):
        """NVML directly (nvidia_ml_py): a sample every 10 ms instead of one nvidia-smi process every 200 ms"""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        reasons_fn = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = ((0x8, 2), (0x40, 3), (0x20, 4), (0x4, 5))        # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self.stop_flag:
            r = int(reasons_fn(h))
            s = [str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx), '', '', '', '']
            for bit, pos in bits:
                s[pos] = 'Active' if r & bit else 'Not Active'
            self.samples.append(s)
            time.sleep(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(self.samples)}


def cpu_port_arm(steps, warmup, cells_per_step=None):
    """The C oracle (a port of the reference's algorithm) on all host threads; one step = `cells` cells x one external step (3000 sf)."""
    from oracle import oracle_py
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.record import make_conf
    oracle_py.build()
    cores = os.cpu_count() or 1
    cells = cells_per_step or 64 * cores
    ctrl = ControllerTables(AGENTS, EXT_AGENT).build()
    conf = make_conf(CONF, **CAPS)
    acts = ext_actions(warmup + steps, cells, seed=1234)
    seeds = np.arange(cells, dtype=np.uint64)
    # the cells must pass through the warm-up steps: time (W + K) steps and W steps, report the difference
    sec, sf, ns = oracle_py.bench(conf, ctrl, seeds, acts, warmup + steps, -1, cores)
    sec_w = 0.0
    if warmup:
        sec_w, sf_w, _ = oracle_py.bench(conf, ctrl, seeds, acts[:warmup], warmup, -1, cores)
        sf -= sf_w
    timed = max(sec - sec_w, 1e-9)
    return {'value': sf / timed, 'seconds': timed, 'subframes': sf, 'cores': cores, 'cells': cells, 'kind': 'port',
            'sample': f'{cells} cells x {steps} external steps of {SF_PER_STEP} subframes (same scenario and policy) on {cores} host threads'}


def cpu_reference_env_arm(steps, warmup, processes=None):
    """The reference's own environment, unmodified, as a multiprocess vector env: one environment per process (its event list is a
    module global), P = host cores, cell g steps the same external actions as cell g of the GPU job. None when oracle/_ref is absent."""
    from oracle import ref_vector_env as rv
    if rv.reference_dir() is None:
        return None
    P = processes or (os.cpu_count() or 1)
    acts = ext_actions(warmup + steps, P, seed=1234)
    r = rv.run(CONF, AGENTS, EXT_AGENT, list(range(P)), acts, warmup, P)
    r['actions'] = acts
    r.update(kind='reference', cells=P,
             sample=f'{P} reference environments (one per process, {r["reference_dir"]}) x {steps} external steps of {SF_PER_STEP} subframes after '
                    f'{warmup} warm-up steps, same scenario, policy and Philox stream as cells 0..{P - 1} of the GPU job')
    return r


CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
CFG3 = {'M': 2000, 'ratio': 1, 'buffer_range': [100, 500], 'reward_criteria': 'average_delay', 'sc_adjustment': True, 'mcs_automatic': False}
CFG4 = {'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True}
NPUSCH_AGENTS = [      # scripts/npusch/experiments_RL.py:40-78
    {'id': 0, 'action_items': ['id', 'Imcs', 'Nrep'], 'obs_items': ['total_ues', 'connection_time', 'loss', 'sinr', 'buffer', 'carrier_state'],
     'next': 1, 'states': ['Scheduling']},
    {'id': 1, 'action_items': ['carrier', 'delay', 'sc'], 'obs_items': [], 'next': -1, 'states': ['Scheduling']},
    {'id': 2, 'action_items': ['carrier', 'ce_level', 'rar_Imcs', 'delay', 'sc', 'Nrep'], 'obs_items': [], 'next': -1, 'states': ['RAR_window']},
    {'id': 3, 'action_items': ['backoff'], 'obs_items': [], 'next': -1, 'states': ['RAR_window_end']},
    {'id': 4, 'action_items': ['th_C1', 'th_C0', 'sc_C0', 'sc_C1', 'sc_C2', 'period_C0', 'period_C1', 'period_C2'], 'obs_items': [],
     'next': -1, 'states': ['NPRACH_update']},
]


def uniform_actions(seed, step, first_cell, n, nvec):
    """uint8[n][len(nvec)], item j uniform on 0..nvec[j]-1: a function of (seed, step, global cell index) alone"""
    rng = np.random.Generator(np.random.Philox(key=[seed, step]))
    r = rng.integers(0, 1 << 32, size=(first_cell + n, len(nvec)), dtype=np.uint64)[first_cell:]
    return ((r * np.asarray(nvec, dtype=np.uint64)) >> 32).astype(np.uint8)


class _Job:
    """what the extra configuration blocks share: device, ranks, reductions (device time = max over ranks, counters = sum)"""

    def __init__(self, dev, rank, world, dist):
        self.dev, self.rank, self.world, self.dist = dev, rank, world, dist

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def reduce(self, ms, sums):
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        v = torch.tensor(list(sums), dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(v, op=self.dist.ReduceOp.SUM)
        return float(t.item()), [float(x) for x in v.tolist()]

    def finish(self, env, name, N, K, ms, d, lbar_sum, carriers, extra):
        """d: this rank's statistics delta over the timed region; returns the block's dict (same on every rank)"""
        from nb_iot_environment_b200.sharding import allreduce_stats
        tot = allreduce_stats(env.stats_tensor().clone())
        ms_job, (sf, ev, lb) = self.reduce(ms, [float(d[13]), float(d[14]), lbar_sum])
        lbar, rho = lb / (K * N * self.world), ev / max(sf, 1.0)
        balg = b_alg(lbar, rho, carriers)
        peak, _ = measured_peak()
        achieved = balg * sf / self.world / (ms_job / 1e3) / 1e9          # per GPU: this rank's kernel against one GPU's HBM
        out = {'workload': name, 'instances_per_gpu': N, 'steps': K, 'ms_per_step': ms_job / K, 'value': sf / (ms_job / 1e3), 'unit': UNIT,
               'faulted_instances': int(tot[16].item()), 'truncated_lists': int(tot[19].item()), 'max_live_ues_per_cell': int(tot[21].item()),
               'max_lookahead_columns': int(tot[20].item()), 'state_GB_per_gpu': env.state.numel() / 1e9,
               'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                            'b_alg_bytes_per_subframe_instance': balg, 'lbar_live_ues': lbar, 'rho_events_per_subframe': rho}}
        out.update(extra)
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        key = extra.get('traffic_key')
        if key and os.path.exists(tp) and key in json.load(open(tp)):
            # DRAM bytes and warp instructions of ONE launch of this block's kernel from ncu (same build, same workload, scaled by the cells of
            # this run); time and counters are this run's
            tj = json.load(open(tp))[key]
            scale = N / float(tj.get('cells', N))
            out['roofline']['traffic'] = tj['dram_bytes_per_launch'] * scale
            wi = float(tj['warp_inst_per_launch']) * scale
            sf_launch = sf / self.world / K
            out['issue'] = {'warp_inst_per_launch': wi, 'inst_per_subframe_instance': wi / max(sf_launch, 1.0), 'inst_per_event': wi / max(ev / self.world / K, 1.0),
                            'ipc_per_sm': wi / ((ms_job / K) * 1e-3 * 1965e6 * 148), 'active_lanes_per_inst': tj.get('active_lanes_per_inst'),
                            'source': tj.get('source')}
        out.pop('traffic_key', None)
        return out


def block_config5(job, local, K, W):
    """BASELINE.json configs[4], the sweep configuration: config-1 traffic (M=1000, ratio 1), the reference's default agents on the device,
    no host round trip; a step = 3000 subframes of every cell in ONE launch (weak scaling: N cells per GPU)."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.sharding import shard_seeds
    N = int(os.environ.get('NBIOT_BENCH_N5', '131072'))
    env = BatchedNBIoTEnv(CFG1, num_envs=N, seeds=shard_seeds(N * job.world, job.rank, job.world), device=local, ue_cap=256, grid_w=1024, hist_cap=512)
    env.reset()
    for s in range(W):
        env.run_until(SF_PER_STEP * (s + 1))
    job.barrier()
    s0 = env.stats_tensor().clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    live = []
    ev[0].record()
    for s in range(K):
        env.run_until(SF_PER_STEP * (W + s + 1))
        ev[s + 1].record()
        live.append(env.stats_tensor()[17:18].clone())
    job.barrier()
    ms = sum(ev[s].elapsed_time(ev[s + 1]) for s in range(K))         # the statistics kernel between two steps is inside (microseconds)
    d = (env.stats_tensor().clone() - s0).cpu().numpy()
    out = job.finish(env, 'configs[4] sweep point: config-1 traffic (M=1000, ratio 1), default agents on the device, 1 step = 3000 subframes per cell in one launch',
                     N, K, ms, d, float(torch.cat(live).double().sum().item()), 1,
                     {'l2': f'state {env.state.numel() / 1e9:.1f} GB per GPU, far larger than L2', 'kernel': env.kernel_name(), 'traffic_key': 'config5'})
    env.close()
    return out


def block_config2(job, local, K, W):
    """BASELINE.json configs[2]: scripts/npusch scenario (scenarios['2000_10_B'], experiments_RL.py:40-78), scheduling agent (UE id, Imcs,
    Nrep) external with uniform-random actions, 65 536 cells per GPU; a step = one external decision of every cell = one launch."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.sharding import shard_seeds
    N = int(os.environ.get('NBIOT_BENCH_N2', '65536'))
    K, W = 10 * K, 50
    env = BatchedNBIoTEnv(CFG3, agents_conf=NPUSCH_AGENTS, ext_agent=0, num_envs=N, seeds=shard_seeds(N * job.world, job.rank, job.world), device=local,
                          ue_cap=CFG3['M'], grid_w=2048, hist_cap=1024)
    # uniform-random actions from a device generator seeded per rank: step s draws the same values whatever --steps is
    g = torch.Generator(device=job.dev); g.manual_seed(7700 + job.rank)
    nvec = torch.tensor(env.action_nvec, device=job.dev)
    acts = (torch.rand((W + K, N, 3), device=job.dev, generator=g) * nvec).to(torch.uint8)
    env.reset()
    for s in range(W):
        env.step(acts[s])
    job.barrier()
    s0 = env.stats_tensor().clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    live = []
    e0.record()
    for s in range(K):
        env.step(acts[W + s])
        if s % 10 == 0:
            live.append(env.stats_tensor()[17:18].clone())
    e1.record()
    job.barrier()
    d = (env.stats_tensor().clone() - s0).cpu().numpy()
    out = job.finish(env, 'configs[2]: NPUSCH control scenario (M=2000, ratio 1, scheduling agent external, uniform-random actions), 1 step = 1 external decision per cell = 1 launch',
                     N, K, e0.elapsed_time(e1), d, float(torch.cat(live).double().sum().item()) * 10.0, 1,
                     {'kernel': 'warp-per-cell (launches of one or two NodeB steps run on it whatever the batch size)'})
    out['instance_decisions_per_s'] = N * job.world * K / (out['ms_per_step'] * K / 1e3)
    env.close()
    return out


def block_config3(job, local, K, W):
    """BASELINE.json configs[3]: two-carrier cell, NPRACH and NPUSCH parameters both chosen from outside at every decision (M=3000, ratio 0.5,
    random traffic), 262 144 cells over the job's GPUs (65 536 on a single GPU); a step = one NodeB decision of every cell = one launch."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    from nb_iot_environment_b200.sharding import shard_seeds
    N = int(os.environ.get('NBIOT_BENCH_N3', '0')) or (65536 if job.world == 1 else 262144 // job.world)
    K, W = 10 * K, 50
    env = BatchedNBIoTEnv(CFG4, num_envs=N, seeds=shard_seeds(N * job.world, job.rank, job.world), n_carriers=2, device=local, all_states_external=True,
                          ue_cap=CFG4['M'], grid_w=2048, hist_cap=2048)
    # uniform over control_max_values for 2 carriers minus what the reference raises on (tools/gen_golden.py policy_action), from a device
    # generator seeded per rank (step s draws the same values whatever --steps is)
    mx = torch.tensor([2, 4, 7, 8, 4, 4, 5, 12, 8, 8, 11, 16, 6, 8, 8, 6, 8, 8, 6, 8, 8, 15, 15], device=job.dev)
    g = torch.Generator(device=job.dev); g.manual_seed(7800 + job.rank)

    def actions(s):
        a = (torch.rand((N, 23), device=job.dev, generator=g) * mx).to(torch.uint8)
        a[:, 5] = torch.where(a[:, 5] == 2, torch.full_like(a[:, 5], 3), a[:, 5])          # sc = 2 is invalid
        rar = env.record_field('state') == 2                                              # RAR_window: ce_level / rar_Imcs share items 1 / 2
        a[:, 1] = torch.where(rar, a[:, 1] % 3, a[:, 1]); a[:, 2] = torch.where(rar, a[:, 2] % 3, a[:, 2])
        return a
    env.reset()
    for s in range(W):
        env.node_step(actions(s))
    job.barrier()
    s0 = env.stats_tensor().clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    live = []
    e0.record()
    for s in range(K):
        env.node_step(actions(W + s))
        if s % 10 == 0:
            live.append(env.stats_tensor()[17:18].clone())
    e1.record()
    job.barrier()
    d = (env.stats_tensor().clone() - s0).cpu().numpy()
    out = job.finish(env, 'configs[3]: two carriers, NPRACH + NPUSCH both controlled from outside at every decision, M=3000 bursty (ratio 0.5, random traffic), '
                          '262144 cells over the GPUs of the job (65536 on one GPU), 1 step = 1 NodeB decision per cell = 1 launch (action upload + state-dependent fix-up inside)',
                     N, K, e0.elapsed_time(e1), d, float(torch.cat(live).double().sum().item()) * 10.0, 2,
                     {'kernel': 'warp-per-cell (launches of one or two NodeB steps run on it whatever the batch size)'})
    out['instance_decisions_per_s'] = N * job.world * K / (out['ms_per_step'] * K / 1e3)
    env.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=12)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--instances', type=int, default=N_PER_GPU, help='instances per GPU')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    K, W = args.steps, max(args.warmup, 0)

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference_env_arm(K, W)
        port = cpu_port_arm(K, W)
        note = ('the unmodified Python reference (system/, controller.py, gym-system/ staged in oracle/_ref), one environment per process, '
                'random generator = the Philox shim; cpu_port = its C restatement (oracle/) on the same cores')
        if r is None:
            r, note = port, 'oracle/_ref is absent on this box: kind=port is the C restatement of the reference (oracle/), validated against golden traces of the unmodified reference'
        line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K, 'warmup': W,
                'ms_per_step': 1e3 * r['seconds'] / max(K, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'int32+f64', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'cpu_cells_per_step': r['cells']},
                'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
                'cpu_port': {'value': port['value'], 'unit': UNIT, 'cores': port['cores'], 'kind': 'port', 'sample': port['sample']},
                'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'note': note}
        print(json.dumps(line))
        return 0

    # anything a library prints on stdout (e.g. NCCL's version banner) goes to stderr: stdout carries ONE JSON line
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import __graft_entry__ as g
    g.build_cuda()
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    W = max(W, 3)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    N = args.instances
    from nb_iot_environment_b200.sharding import shard_seeds, allreduce_stats
    seeds = shard_seeds(N * world, rank, world)              # instance g of the job has Philox key g on any number of GPUs
    if os.environ.get('NBIOT_BENCH_SAME'):                    # measurement knob: identical cells (upper bound of code sharing between warps)
        seeds = np.zeros_like(np.asarray(seeds))
    env = BatchedNBIoTEnv(CONF, agents_conf=AGENTS, ext_agent=EXT_AGENT, num_envs=N, seeds=seeds, device=local_rank, **CAPS)
    total_steps = W + 2 * K
    acts_host = ext_actions(total_steps, N, seed=1234, first_cell=rank * N)
    if os.environ.get('NBIOT_BENCH_SAME'):
        acts_host[:] = acts_host[:, :1]
    acts_pinned = torch.from_numpy(acts_host).pin_memory()
    acts_dev = acts_pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    env.reset()
    torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up
    for s in range(W):
        env.step(acts_dev[s])
    barrier()
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- timed region 1: device-resident inputs, CUDA events on the launching stream, L2 flushed between steps
    import ctypes
    from nb_iot_environment_b200 import _cabi
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    live, stats_prev = [], env.stats_tensor().clone()
    t_wall0 = time.perf_counter()
    for s in range(K):
        flush.fill_(s & 0xFF)
        a = acts_dev[W + s]
        st = torch.cuda.current_stream(dev).cuda_stream
        ev[s][0].record()
        _cabi.check(env.L.nbiot_advance(env._h, a.data_ptr(), -1, 1 << 30, st))
        ev[s][1].record()
        _cabi.check(env.L.nbiot_gather_obs(env._h, env._items, env._n_items, env._obs.data_ptr(), env._reward.data_ptr(), st))
        ev[s][2].record()
        live.append(env.stats_tensor()[17:18].clone())
    stats_now = env.stats_tensor().clone()
    job_stats = allreduce_stats(stats_now.clone())                       # the one collective of the path: episode statistics (NCCL)
    barrier()
    wall1 = time.perf_counter() - t_wall0
    sim_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(K)]
    step_ms = [ev[s][0].elapsed_time(ev[s][2]) for s in range(K)]
    d = (stats_now - stats_prev).cpu().numpy()
    subframes, events = int(d[13]), int(d[14])
    lbar = float(torch.cat(live).double().mean().item()) / N
    faulted = int(stats_now[16].item())

    # ---- timed region 2: end to end through the host-buffer C-ABI call
    torch.cuda.synchronize(dev)
    obs_host = np.empty((N, max(1, env.obs_dim)), dtype=np.float64)
    rew_host = np.empty(N, dtype=np.float64)
    e2e_prev = env.stats_tensor().clone()
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        env.step_host(acts_host[W + K + s], obs_host, rew_host)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_now = env.stats_tensor().clone()
    e2e_subframes = int((e2e_now - e2e_prev)[13].item())                  # what the cells really advanced (a faulted cell stands still)
    launches = env.launch_count - launches0
    final_stats = allreduce_stats(e2e_now.clone())
    geometry, kernel_name, obs_dim = env.launch_geometry(), env.kernel_name(), env.obs_dim

    # ---- warp-busy fraction of one launch (per-cell start / end times, NBIOT_PROF_INST): rank 0, a separate small leg
    busy = None
    if rank == 0 and not os.environ.get('NBIOT_BENCH_NO_BUSY'):
        try:
            busy = warp_busy_fraction(local_rank, N, acts_host[:W + 1])
        except Exception as e:                                            # a measurement aid must not fail the benchmark
            print('warp-busy leg failed:', repr(e), file=sys.stderr)
    env.close()
    del env
    torch.cuda.empty_cache()

    # ---- the other configurations of BASELINE.json, measured in the same run (every rank takes part: their timing is max over ranks)
    job = _Job(dev, rank, world, dist)
    blocks = {}
    want = os.environ.get('NBIOT_BENCH_BLOCKS', '5,2,3').split(',')
    for key, fn in (('5', block_config5), ('2', block_config2), ('3', block_config3)):
        if key in want:
            blocks['config' + key] = fn(job, local_rank, K, W)
            torch.cuda.empty_cache()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    dev_s = sum(step_ms) / 1e3
    t = torch.tensor([dev_s, e2e_s, float(subframes), sum(sim_ms) / 1e3, float(e2e_subframes)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s, e2e_s, sim_s = float(tmax[0]), float(tmax[1]), float(tmax[3])
        subframes_all, e2e_subframes_all = float(tsum[2]), float(tsum[4])
    else:
        sim_s, subframes_all, e2e_subframes_all = sum(sim_ms) / 1e3, float(subframes), float(e2e_subframes)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = subframes_all / dev_s
    e2e_value = e2e_subframes_all / e2e_s
    rho = events / max(subframes, 1)
    balg = b_alg(lbar, rho, 1)
    peak, peak_src = measured_peak()
    kernel_ms = sum(sim_ms) / K
    achieved = balg * (subframes / K) / (kernel_ms / 1e3) / 1e9
    clocks = sampler.summary()
    traffic, issue = None, None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get('dram_bytes_per_launch')
        if tj.get('warp_inst_per_launch'):
            # instructions per launch come from ncu (smsp__inst_executed.sum of the same build and workload, profiles/traffic.json names the
            # report); time, clock, events and subframes are this run's
            wi, mhz, sms = float(tj['warp_inst_per_launch']), float(clocks.get('sm_mhz') or 1965.0), 148
            issue = {'warp_inst_per_launch': wi, 'inst_per_subframe_instance': wi / (subframes / K), 'inst_per_event': wi / max(events / K, 1),
                     'ipc_per_sm': wi / (kernel_ms * 1e-3 * mhz * 1e6 * sms), 'ipc_frac_of_4': wi / (kernel_ms * 1e-3 * mhz * 1e6 * sms) / 4.0,
                     'active_lanes_per_inst': tj.get('active_lanes_per_inst'), 'warp_busy_fraction': busy,
                     'source': 'warp_inst_per_launch, active_lanes_per_inst: ncu (' + str(tj.get('source')) + '); kernel time, clock, events: this run; '
                               'warp_busy_fraction: per-cell globaltimer stamps of one launch of this run'}
    bad = {'headline': (int(final_stats[16].item()), int(final_stats[19].item()))}
    bad.update({k: (b['faulted_instances'], b['truncated_lists']) for k, b in blocks.items()})
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': 1e3 * dev_s / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'int32+f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'instances_per_gpu': N, 'subframes_per_step': SF_PER_STEP, 'l2': 'flushed between timed steps (256 MiB write)',
                   'caps': CAPS, 'timing': 'CUDA events on the launching stream; max over ranks of the summed step times',
                   'actions': 'function of (seed 1234, step, global cell index): identical for any --steps / --warmup / --gpus',
                   'faulted_instances': int(final_stats[16].item()), 'truncated_lists': int(final_stats[19].item()),
                   'max_lookahead_columns': int(final_stats[20].item()), 'max_live_ues_per_cell': int(final_stats[21].item()),
                   'launch_geometry': geometry, 'kernel': kernel_name},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(N * 8), 'd2h_bytes_per_step': int(N * obs_dim * 8 + N * 8),
                'ms_per_step': 1e3 * e2e_s / K, 'api': 'BatchedNBIoTEnv.step_host -> nbiot_step_host (pinned staging, synchronous)'},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'peak_source': peak_src, 'kernel': 'nbiot_sim_kernel', 'kernel_ms': kernel_ms,
                     'b_alg_bytes_per_subframe_instance': balg, 'lbar_live_ues': lbar, 'rho_events_per_subframe': rho,
                     'note': 'algorithmic bytes of SURVEY 8(d) (a per-subframe UE scan); the kernel skips idle subframes, so its DRAM traffic is far below this: '
                             'see `issue` for the ceiling it is actually under'},
        'issue': issue,
        'wall_s_timed_region': wall1,
    }
    line.update(blocks)
    print('episode statistics (job):', [int(v) for v in job_stats.tolist()], file=sys.stderr)     # equal for every kernel / scheduler that runs the same batch
    if world == 1 and not os.environ.get('NBIOT_BENCH_NO_CPU'):          # (measurement sweeps skip the CPU legs)
        port = cpu_port_arm(40, 0, cells_per_step=64 * (os.cpu_count() or 1))       # ~15 CPU-seconds of work
        line['cpu_port'] = {'value': port['value'], 'unit': UNIT, 'cores': port['cores'], 'kind': 'port', 'sample': port['sample']}
        r = None
        try:
            r = cpu_reference_env_arm(10, 3)                              # ~1 s per process + import
        except Exception as e:
            print('reference environment leg failed:', repr(e), file=sys.stderr)
        if r is not None:
            try:
                line['parity_vs_reference'] = reference_parity(local_rank, r)
            except Exception as e:
                print('live parity leg failed:', repr(e), file=sys.stderr)
        r = r or port
        line['cpu_baseline'] = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()
    if any(f or tr for f, tr in bad.values()):
        print('FAULTED OR TRUNCATED INSTANCES in the timed workload (faulted, truncated):', bad, file=sys.stderr)
        return 3
    return 0


def reference_parity(local_rank, r):
    """The cells the reference environments just simulated on the host cores, simulated again on the GPU with the same seeds and
    actions: observation and reward of every step against the UNMODIFIED reference, live on this box."""
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv
    acts, traces = r['actions'], r['traces']
    P = acts.shape[1]
    env = BatchedNBIoTEnv(CONF, agents_conf=AGENTS, ext_agent=EXT_AGENT, num_envs=P, seeds=list(range(P)), device=local_rank, **CAPS)
    env.reset()
    worst, rewards_equal = 0.0, True
    for s in range(acts.shape[0]):
        obs, rew, _, _, _ = env.step(acts[s])
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        for i in range(P):
            o_ref, r_ref = np.asarray(traces[i][s][0], dtype=np.float64), traces[i][s][1]
            den = np.maximum(np.abs(o_ref), 1e-300)
            worst = max(worst, float(np.max(np.where(o_ref == obs[i], 0.0, np.abs(obs[i] - o_ref) / den))))
            rewards_equal = rewards_equal and (r_ref == rew[i])
    env.close()
    return {'cells': P, 'steps': int(acts.shape[0]), 'obs_max_rel_err': worst, 'rewards_equal': bool(rewards_equal),
            'against': 'the unmodified reference environments of cpu_baseline (same seeds, same actions), every step'}


def warp_busy_fraction(local_rank, N, acts):
    """mean over cells of (cell busy time) / (launch duration x cells that can run at once ... ) -- here simply the mean busy time of a
    cell over the span of the launch, for one step of the headline workload (the one-wave tail)"""
    import ctypes
    import torch
    from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi
    old = os.environ.get('NBIOT_PROF_INST')
    os.environ['NBIOT_PROF_INST'] = '1'
    try:
        env = BatchedNBIoTEnv(CONF, agents_conf=AGENTS, ext_agent=EXT_AGENT, num_envs=N, device=local_rank, **CAPS)
    finally:
        if old is None:
            del os.environ['NBIOT_PROF_INST']
        else:
            os.environ['NBIOT_PROF_INST'] = old
    env.reset()
    for a in acts:
        env.step(torch.from_numpy(a).to(env.device))
    torch.cuda.synchronize(env.device)
    out = np.zeros((N, 4), dtype=np.uint64)
    _cabi.check(env.L.nbiot_instance_times(env._h, out.ctypes.data))
    env.close()
    t0, t1 = out[:, 0].astype(np.float64), out[:, 1].astype(np.float64)
    span = t1.max() - t0.min()
    return float((t1 - t0).mean() / span) if span > 0 else None


if __name__ == '__main__':
    sys.exit(main())
