#!/usr/bin/env python3
"""Ad-hoc: step the Python reference and the C oracle side by side and report the first difference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.refharness import RefCell
from oracle.oracle_py import OracleCell
from nb_iot_environment_b200.record import RECORD_DTYPE, FLOAT_FIELDS, INT_FIELDS, NON_PARITY_FIELDS

DEFAULT_ACTION = [0, 1, 1, 4, 0, 3, 1, 3, 6, 2, 4, 12, 3, 0, 1, 2, 0, 1, 3, 2, 1, 8, 11]


def diff(a, b, rtol=1e-9):
    out = []
    for n in INT_FIELDS:
        if n in NON_PARITY_FIELDS:
            continue
        if not np.array_equal(a[n], b[n]):
            out.append((n, a[n], b[n]))
    for n in FLOAT_FIELDS:
        if not np.allclose(a[n], b[n], rtol=rtol, atol=1e-12):
            out.append((n, a[n], b[n]))
    return out


RANDOM_POLICY = os.environ.get('POLICY', 'fixed') == 'random'
prng = np.random.default_rng(12345)


def policy(rec, prng, nc=1):
    # uniform over control_max_values (parameters.py:191-218) minus the values the reference raises on
    nc = NC
    mx = [nc - 1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, 3 if nc == 1 else 7, 5, 7, 3 if nc == 1 else 7, 5, 7, 3 if nc == 1 else 7, 14, 14]
    a = [int(prng.integers(0, m + 1)) for m in mx]
    if a[5] == 2:
        a[5] = 3
    if rec['state'] == 2:      # RAR_window: index 1 = ce_level (0..2), index 2 = rar_Imcs (0..2)
        a[1] = int(prng.integers(0, 3)); a[2] = int(prng.integers(0, 3))
    return a


NC = 1


def main():
    global NC
    conf = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    nc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    NC = nc
    if len(sys.argv) > 4:
        conf.update(eval(sys.argv[4]))
    ref = RefCell(conf, seed, n_carriers=nc, log_events=True)
    orc = OracleCell(conf, seed, n_carriers=nc)
    orc.enable_event_log()
    a, b = ref.reset(), orc.reset()
    n = 0
    t0 = time.time()
    while True:
        d = diff(a, b)
        if d:
            print(f'MISMATCH at decision {n}, ref time {a["time"]} state {a["state"]}, oracle time {b["time"]} state {b["state"]}')
            for name, x, y in d[:12]:
                print('  ', name, '\n     ref:', x, '\n     orc:', y)
            ev_r = np.array(ref.events); ev_o = orc.event_log()
            m = min(len(ev_r), len(ev_o))
            bad = np.nonzero((ev_r[:m] != ev_o[:m]).any(axis=1))[0]
            if len(bad):
                i = bad[0]; print('first event divergence at pop', i, 'ref', ev_r[max(0,i-5):i+3].tolist(), 'orc', ev_o[max(0,i-5):i+3].tolist())
            else:
                print('event logs agree for', m, 'pops; lens', len(ev_r), len(ev_o))
            return 1
        if a['time'] >= T:
            break
        act = policy(a, prng) if RANDOM_POLICY else DEFAULT_ACTION
        a, b = ref.step(act), orc.step(act)
        n += 1
    print(f'OK: {n} decisions to t={a["time"]}, draws={ref.rng.ctr}, ref counters {ref.counters()}')
    co = orc.counters()
    cr = ref.counters()
    for k in cr:
        if k == 'av_delay':
            assert abs(cr[k] - co[k]) <= 1e-9 * max(1, abs(cr[k])), (k, cr[k], co[k])
        else:
            assert cr[k] == co[k], (k, cr[k], co[k])
    print('counters equal; pops', co['pops'], len(ref.events), 'elapsed', time.time() - t0)
    return 0


if __name__ == '__main__':
    sys.exit(main())
