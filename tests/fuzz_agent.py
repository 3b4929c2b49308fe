"""Seeded random inputs of the model-based NPRACH agent (agent_nprach.NPRACH_THAgent.get_action): per step the detection / collision lists of the
NPRACH occasions of an update period, the loss distribution, the number of incoming samples and the margin. Loads range from idle cells to far
beyond what a cell produces, so that every branch of get_action (gathering, table look-up below / inside / above the rate range, no_th) is taken."""
import numpy as np


def agent_sequence(k, steps=60):
    r = np.random.default_rng(60_000 + k)
    no_th = bool(r.integers(0, 2))
    scale = float(r.choice([0.2, 1.0, 3.0, 8.0]))
    seq = []
    for s in range(steps):
        det, col = [], []
        for ce in range(3):
            n = int(r.integers(0, 49)) if r.integers(0, 8) else 0
            lam = scale * float(r.uniform(0.0, 6.0))
            d = np.minimum(r.poisson(lam, size=n), 48)
            c = np.minimum(r.poisson(lam / 3.0, size=n), 47)
            det.append([int(x) for x in d]); col.append([int(x) for x in c])
        dist = r.dirichlet(np.ones(15) * float(r.choice([0.3, 1.0, 5.0])))
        seq.append({'NPRACH_detection': det, 'NPRACH_collision': col, 'distribution': [float(x) for x in dist],
                    'incoming': [0.0] * int(r.integers(0, 40)), 'margin': float(r.uniform(0.3, 3.4)),
                    'departures': int(r.integers(0, 50)), 'NPRACH_occupation': float(r.uniform(0, 1)), 'service_times': [int(x) for x in r.integers(1, 500, size=int(r.integers(0, 6)))]})
    return no_th, seq
