#!/usr/bin/env python3
"""bench.py -- throughput of the batched NB-IoT cell simulator on the reference's NPRACH-control scenario.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): scripts/nprach of the reference -- M=1000 devices per cell, ratio 0.5
('mixed_05', evaluate_RL.py:30-34,108-114), agents of evaluate_RL.py:46-77 with the NPRACH agent (id 3) external,
4096 cell instances per GPU. A STEP is one external decision of every instance = 3000 simulated subframes per
instance (the NPRACH update period). External actions are uniform-random feasible NPRACH configurations
(threshold pair out of the 105 of parameters.th_pairs, per-CE subcarrier and period indices), pre-generated on the
host from a seeded generator.

One JSON line (rank 0): `value` = simulated subframe-instances/s with inputs resident in HBM, timed with CUDA
events on the launching stream, L2 flushed between timed steps; `e2e` = the same through the host-buffer C-ABI
call (actions host->device, observation + reward device->host every step); `roofline` = algorithmic bytes of
SURVEY.md 8(d) (constants in include/nbiot_roofline.h, L-bar and rho_ev measured on the device) over the sim
kernel's CUDA-event time, against the measured HBM peak; `cpu_baseline` = the CPU oracle (a C port of the
reference's algorithm) on this box's host cores on a bounded sample. `--impl reference` prints the CPU arm alone.
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


def cpu_reference_arm(steps, warmup, cells_per_step=None):
    """The CPU arm: the oracle port on all host threads; one step = `cells` cells x one external step (3000 sf)."""
    from oracle import oracle_py
    from nb_iot_environment_b200.controller import ControllerTables
    from nb_iot_environment_b200.record import make_conf
    oracle_py.build()
    cores = os.cpu_count() or 1
    cells = cells_per_step or 64 * cores
    ctrl = ControllerTables(AGENTS, EXT_AGENT).build()
    conf = make_conf(CONF, **CAPS)
    acts = ext_actions(warmup + steps, cells, seed=999)
    seeds = np.arange(cells, dtype=np.uint64)
    # the cells must pass through the warm-up steps: time (W + K) steps and W steps, report the difference
    sec, sf, ns = oracle_py.bench(conf, ctrl, seeds, acts, warmup + steps, -1, cores)
    sec_w = 0.0
    if warmup:
        sec_w, sf_w, _ = oracle_py.bench(conf, ctrl, seeds, acts[:warmup], warmup, -1, cores)
        sf -= sf_w
    timed = max(sec - sec_w, 1e-9)
    return {'value': sf / timed, 'seconds': timed, 'subframes': sf, 'cores': cores, 'cells': cells,
            'sample': f'{cells} cells x {steps} external steps of {SF_PER_STEP} subframes (same scenario and policy) on {cores} host threads'}


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
        r = cpu_reference_arm(K, W)
        line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K, 'warmup': W,
                'ms_per_step': 1e3 * r['seconds'] / max(K, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'int32+f64', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'cpu_cells_per_step': r['cells']},
                'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
                'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'note': 'the reference is pure Python and is not present on this box; kind=port is its C restatement (oracle/), '
                        'validated against golden traces of the unmodified reference'}
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
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        env.step_host(acts_host[W + K + s], obs_host, rew_host)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = env.launch_count - launches0

    dev_s = sum(step_ms) / 1e3
    t = torch.tensor([dev_s, e2e_s, float(subframes), sum(sim_ms) / 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s, e2e_s, sim_s = float(tmax[0]), float(tmax[1]), float(tmax[3])
        subframes_all = float(tsum[2])
    else:
        sim_s, subframes_all = sum(sim_ms) / 1e3, float(subframes)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = subframes_all / dev_s
    e2e_value = (N * SF_PER_STEP * K * world) / e2e_s
    rho = events / max(subframes, 1)
    balg = b_alg(lbar, rho, 1)
    peak, peak_src = measured_peak()
    achieved = balg * (subframes / K) / (sum(sim_ms) / K / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get('dram_bytes_per_launch')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': 1e3 * dev_s / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'int32+f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'instances_per_gpu': N, 'subframes_per_step': SF_PER_STEP, 'l2': 'flushed between timed steps (256 MiB write)',
                   'caps': CAPS, 'timing': 'CUDA events on the launching stream; max over ranks of the summed step times',
                   'faulted_instances': int(job_stats[16].item()), 'max_lookahead_columns': int(job_stats[20].item()),
                   'max_live_ues_per_cell': int(job_stats[21].item()), 'launch_geometry': env.launch_geometry()},
        'clocks': sampler.summary(),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(N * 8), 'd2h_bytes_per_step': int(N * env.obs_dim * 8 + N * 8),
                'ms_per_step': 1e3 * e2e_s / K, 'api': 'BatchedNBIoTEnv.step_host -> nbiot_step_host (pinned staging, synchronous)'},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'peak_source': peak_src, 'kernel': 'nbiot_sim_kernel', 'kernel_ms': sum(sim_ms) / K,
                     'b_alg_bytes_per_subframe_instance': balg, 'lbar_live_ues': lbar, 'rho_events_per_subframe': rho,
                     'note': 'algorithmic bytes of SURVEY 8(d) (a per-subframe UE scan); the kernel skips idle subframes, so its DRAM traffic is far below this'},
        'wall_s_timed_region': wall1,
    }
    print('episode statistics (job):', [int(v) for v in job_stats.tolist()], file=sys.stderr)     # equal for every kernel / scheduler that runs the same batch
    if os.environ.get('NBIOT_SCHED', '').startswith('queue'):
        qs = (ctypes.c_ulonglong * 32)()
        _cabi.check(env.L.nbiot_queue_stats(env._h, qs))
        line['config']['scheduler'] = 'migrating instances (NBIOT_SCHED=queue)'
        print('queue stats: tasks', list(qs)[0:10], 'cycles', list(qs)[10:20], 'hops', qs[20], 'local', qs[21], 'switches', qs[22],
              'idle_polls', qs[23], 'bail', qs[24], file=sys.stderr)
    if world == 1 and not os.environ.get('NBIOT_BENCH_NO_CPU'):          # (measurement sweeps skip the CPU leg)
        r = cpu_reference_arm(40, 0, cells_per_step=64 * (os.cpu_count() or 1))     # ~15 CPU-seconds of work
        line['cpu_baseline'] = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
