#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED Python reference (/root/reference)
with the Philox shim injected as its random generator (tools/refharness.py).

Run in the build container only:   python tools/gen_golden.py [case ...]

Every fixture holds, for each seed of the case, the action applied at every decision and the
complete decision record (reward + info dict, include/nbiot_record.h) the reference returned,
plus the per-period lists of the NPRACH_update flavour and the final PerfMonitor counters.
The reference ships no tests or golden vectors of its own, so these
self-generated traces are what pins the CPU oracle and, through it, the CUDA path.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')

DEFAULT_ACTION = [0, 1, 1, 4, 0, 3, 1, 3, 6, 2, 4, 12, 3, 0, 1, 2, 0, 1, 3, 2, 1, 8, 11]

CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}

CASES = {
    # config 1: the shipped env, fixed default action (notebook cell 8 loop)
    'cfg1_default': dict(conf=CFG1, nc=1, policy='fixed', seeds=[0, 1, 2, 3], decisions=1400),
    # config 2 traffic variants of scripts/nprach/evaluate_RL.py:30-34, fixed action
    'cfg2_mixed_random_traffic': dict(conf=dict(CFG1, ratio=0.5, random=True), nc=1, policy='fixed', seeds=[10, 11], decisions=1400),
    # every control item uniformly random at every decision (exercises the whole action decode)
    'cfg2_random_policy': dict(conf=dict(CFG1, ratio=0.5), nc=1, policy='random', seeds=[20, 21, 22], decisions=1400),
    # config 3: scenarios['2000_10_B'] (scenarios.py:14-24), agent-chosen MCS / repetitions
    'cfg3_npusch_random_policy': dict(conf={'M': 2000, 'ratio': 1, 'buffer_range': [100, 500], 'reward_criteria': 'average_delay',
                                            'statistics': False, 'sc_adjustment': True, 'mcs_automatic': False},
                                      nc=1, policy='random', seeds=[30, 31, 32], decisions=1400),
    # the non-default branches of step_data / step_RAR_window / get_node_info
    'manual_branches': dict(conf={'M': 2000, 'ratio': 1, 'buffer_range': [100, 600], 'reward_criteria': 'log_invd_users',
                                  'mcs_automatic': False, 'tx_all_buffer': False, 'ce_mcs_automatic': False,
                                  'sc_adjustment': False, 'sort_criterium': 'loss'},
                            nc=1, policy='random', seeds=[40, 41], decisions=1400),
    # config 4: two carriers, bursty arrivals, everything agent-controlled
    'cfg4_two_carriers_bursty': dict(conf={'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True},
                                     nc=2, policy='random', seeds=[50, 51, 52], decisions=1400),
    'cfg4_two_carriers_fixed': dict(conf={'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'users'},
                                    nc=2, policy='fixed', seeds=[55], decisions=1400),
    # RA congestion: thousands of contenders per NPRACH, backoff draws dominate
    'overload': dict(conf={'M': 20000, 'ratio': 0.2, 'buffer_range': [256, 256], 'reward_criteria': 'accumulated_delay'},
                     nc=1, policy='fixed', seeds=[60], decisions=700),
    # clustered placement (population.py:75-101) and the statistics histories (perf_monitor.py:61-84, :161-171)
    'clustered_statistics': dict(conf=dict(CFG1, ratio=0.5, clustered=True, statistics=True), nc=1, policy='fixed', seeds=[80, 81], decisions=1400),
    'clustered_tight_small_cell': dict(conf=dict(CFG1, clustered=True, cluster_ues=[3, 8], cluster_prob=0.9, cluster_range=2.0, max_d=5.0, statistics=True),
                                       nc=1, policy='random', seeds=[85], decisions=1400),
    # SURVEY 8(f) rank 3, the rest: PerfMonitor's `traces` lists and the remaining `statistics` lists (perf_monitor.py:61-79, :111-140, :185-202)
    'traces_statistics': dict(conf=dict(CFG1, ratio=0.5, traces=True, statistics=True), nc=1, policy='random', seeds=[90, 91], decisions=1200),
    'traces_two_carriers': dict(conf={'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True,
                                      'traces': True, 'statistics': True}, nc=2, policy='random', seeds=[95], decisions=800),
    # Channel(capture_effect=True) (channel.py:158-159, :200-219): CAPTURE / COLLIDED states, msg3 grants with several contenders
    'capture_effect': dict(conf={'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'traces': True}, nc=1,
                           policy='random', seeds=[97, 98], decisions=1400, capture=True),
    'capture_effect_two_carriers': dict(conf={'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'users', 'random': True, 'traces': True},
                                        nc=2, policy='random', seeds=[99], decisions=900, capture=True),
    'invd_reward_small_cell': dict(conf=dict(CFG1, reward_criteria='invd_users', max_d=3.0, M=600), nc=1, policy='random', seeds=[70], decisions=800),
}


def policy_action(kind, rec, prng, nc):
    if kind == 'fixed':
        return list(DEFAULT_ACTION)
    # uniform over control_max_values (parameters.py:191-218) minus what the reference raises on
    sc = 3 if nc == 1 else 7
    mx = [nc - 1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, sc, 5, 7, sc, 5, 7, sc, 14, 14]
    a = [int(prng.integers(0, m + 1)) for m in mx]
    if a[5] == 2:
        a[5] = 3
    if rec['state'] == 2:   # RAR_window: ce_level / rar_Imcs share indices 1 / 2
        a[1] = int(prng.integers(0, 3))
        a[2] = int(prng.integers(0, 3))
    return a


def run_case(name):
    from tools.refharness import RefCell
    from nb_iot_environment_b200.record import RECORD_DTYPE
    case = CASES[name]
    S, D = len(case['seeds']), case['decisions']
    records = np.zeros((S, D + 1), dtype=RECORD_DTYPE)
    actions = np.zeros((S, D, 23), dtype=np.int8)
    counters, hists, events, stats, pmlists = [], [], [], [], []
    for si, seed in enumerate(case['seeds']):
        cell = RefCell(case['conf'], seed, n_carriers=case['nc'], log_events=True, capture_effect=case.get('capture', False))
        prng = np.random.default_rng(1000 + seed)
        rec = cell.reset()
        records[si, 0] = rec
        for d in range(D):
            a = policy_action(case['policy'], rec, prng, case['nc'])
            actions[si, d] = a
            rec = cell.step(a)
            records[si, d + 1] = rec
        counters.append(cell.counters())
        if case['conf'].get('statistics'):
            pm = cell.perf
            stats.append((np.asarray(pm.delay_history, np.int32), np.asarray(pm.connection_history, np.int32),
                          np.asarray(pm.access_history, np.int32), np.asarray(pm.th_history, np.float64)))
        pl = {}
        pm = cell.perf
        if case['conf'].get('traces'):
            for k in ('rar_detections', 'rar_errors', 'rar_collisions', 'nprach_arrivals', 'nprach_detections'):
                for ce in range(3):
                    pl[f'{k}_{ce}'] = np.asarray(getattr(pm, k)[ce], dtype=np.int64).reshape(-1, 2)
            for k in ('t_rar_detections', 't_nprach_arrivals', 't_nprach_detections'):
                pl[k] = np.asarray(getattr(pm, k), dtype=np.int64).reshape(-1, 2)
        if case['conf'].get('statistics'):
            for k in ('RAR_in', 'RAR_attmp', 'RAR_sent', 'RAR_detected'):
                for ce in range(3):
                    pl[f'{k}_{ce}'] = np.asarray(getattr(pm, k)[ce], dtype=np.int64)
            for ce in range(3):
                pl[f'npusch_delivered_{ce}'] = np.asarray(pm.npusch_delivered[ce], dtype=np.int64).reshape(-1, 2)
            for k in ('msg3_resources_history', 'NPUSCH_resources_history', 'NPRACH_resources_history'):
                pl[k] = np.asarray(getattr(pm, k), dtype=np.int64)
        pmlists.append(pl)
        hists.append(cell.hists)
        events.append(np.asarray(cell.events, dtype=np.int32))
        cell.close()
        print(f'  {name} seed {seed}: t={rec["time"]} draws={cell.rng.ctr} pops={len(cell.events)}', flush=True)
    # flatten the per-period lists: one row of offsets per seed
    flat = {}
    for si in range(S):
        inc = [h[0] for h in hists[si]]
        dl = [h[1] for h in hists[si]]
        sv = [h[2] for h in hists[si]]
        flat[f'incoming_{si}'] = np.concatenate(inc) if inc else np.zeros(0)
        flat[f'incoming_len_{si}'] = np.asarray([len(x) for x in inc], dtype=np.int32)
        flat[f'delays_{si}'] = np.concatenate(dl) if dl else np.zeros(0, dtype=np.int32)
        flat[f'service_{si}'] = np.concatenate(sv) if sv else np.zeros(0, dtype=np.int32)
        flat[f'delays_len_{si}'] = np.asarray([len(x) for x in dl], dtype=np.int32)
        flat[f'events_{si}'] = events[si][:20000]
        if stats:
            flat[f'stat_delay_{si}'], flat[f'stat_conn_{si}'], flat[f'stat_access_{si}'], flat[f'stat_th_{si}'] = stats[si]
        for k, v in pmlists[si].items():
            flat[f'pm_{k}_{si}'] = v
    meta = {'name': name, 'conf': case['conf'], 'n_carriers': case['nc'], 'seeds': case['seeds'], 'policy': case['policy'],
            'decisions': D, 'counters': counters, 'capture_effect': bool(case.get('capture', False)), 'record_itemsize': RECORD_DTYPE.itemsize,
            'record_descr': str(RECORD_DTYPE.descr)}
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), records=records.view(np.uint8).reshape(S, D + 1, -1),
                        actions=actions, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **flat)


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get('NBIOT_GOLDEN_CHILD') == '1':
        run_case(names[0])
        return 0
    # one process per case: the reference fixes its carrier count (and its FEL) at import time
    for n in names:
        print('case', n, flush=True)
        env = dict(os.environ, NBIOT_GOLDEN_CHILD='1')
        subprocess.check_call([sys.executable, os.path.abspath(__file__), n], env=env)
    return 0


if __name__ == '__main__':
    sys.exit(main())
