"""Seeded random configurations and a random NodeB-level policy for the fuzz tests: every conf key of the reference's
create_system (system_creator.py:22-44) that the batched path accepts, one or two carriers, capture effect on / off."""
import numpy as np

REWARDS = ['accumulated_delay', 'average_delay', 'throughput', 'users', 'invd_users', 'log_invd_users']     # perf_monitor.py:29-37 minus the one the reference fails on


def fuzz_case(k):
    """-> (conf, n_carriers, capture_effect)"""
    r = np.random.default_rng(9000 + k)
    lo = int(r.integers(20, 300))
    conf = {
        'M': int(r.integers(150, 2200)),
        'ratio': float(r.choice([0.2, 0.5, 0.8, 1.0])),
        'buffer_range': [lo, lo + int(r.integers(10, 500))],
        'reward_criteria': str(r.choice(REWARDS)),
        'sort_criterium': str(r.choice(['t_connection', 'loss'])),
        'sc_adjustment': bool(r.integers(0, 2)),
        'mcs_automatic': bool(r.integers(0, 2)),
        'ce_mcs_automatic': bool(r.integers(0, 2)),
        'tx_all_buffer': bool(r.integers(0, 2)),
        'random': bool(r.integers(0, 2)),
    }
    if r.integers(0, 3) == 0:
        conf['max_d'] = float(r.choice([0.4, 1.5, 4.0]))
    if r.integers(0, 3) == 0:
        a = int(r.integers(2, 12))
        conf.update(clustered=True, cluster_ues=[a, a + int(r.integers(1, 20))], cluster_prob=float(r.choice([0.2, 0.5, 0.9])), cluster_range=float(r.choice([0.1, 0.5, 1.0])))
    if r.integers(0, 4) == 0:
        conf['statistics'] = True
    if r.integers(0, 4) == 0:
        conf['traces'] = True
    return conf, int(r.integers(1, 3)), bool(r.integers(0, 3) == 0)


def random_action(state, prng, nc):
    """uniform over control_max_values (parameters.py:191-218) minus the values the reference raises on (tools/compare_ref_oracle.py)"""
    mx = [nc - 1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, 3 if nc == 1 else 7, 5, 7, 3 if nc == 1 else 7, 5, 7, 3 if nc == 1 else 7, 14, 14]
    a = [int(prng.integers(0, m + 1)) for m in mx]
    if a[5] == 2:
        a[5] = 3
    if state == 2:             # RAR_window: index 1 = ce_level (0..2), index 2 = rar_Imcs (0..2)
        a[1] = int(prng.integers(0, 3)); a[2] = int(prng.integers(0, 3))
    return a


def controller_case(k):
    """-> (conf, n_carriers, capture_effect, agents_conf, ext_agent, fixed actions per agent, rng for the external actions):
    one of the agent sets of the reference's scripts, a random external agent, random fixed actions for the others"""
    from nb_iot_environment_b200 import parameters as par
    from test_gpu_parity import NPRACH_AGENTS, NPUSCH_AGENTS
    conf, nc, cap = fuzz_case(k)
    r = np.random.default_rng(50_000 + k)
    agents = [par.agents_conf, NPRACH_AGENTS, NPUSCH_AGENTS][int(r.integers(0, 3))]
    ext = int(r.integers(0, len(agents)))

    def draw(name):
        v = int(r.integers(0, par.control_max_values[name][nc - 1] + 1))
        if name == 'backoff' and v == 12:
            v = 11                                         # control_max_values allows 12, backoff_list has 12 entries: the reference raises IndexError
        return 3 if name == 'sc' and v == 2 else v        # sc index 2 is the value the reference raises on (parameters.py:340)
    fixed = [{n: draw(n) for n in a['action_items']} for a in agents]
    for f in fixed:                                        # keep th_C1 <= th_C0 half of the time (both orders are legal)
        if 'th_C1' in f and r.integers(0, 2):
            f['th_C1'] = min(f['th_C1'], f['th_C0'])
    return conf, nc, cap, agents, ext, fixed, draw
