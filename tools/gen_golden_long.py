#!/usr/bin/env python3
"""BASELINE.json configs[0] as a compact golden: the shipped single-cell env (M=1000, ratio=1, buffer [100,600], default
action vector, notebook cell 8 loop) run by the UNMODIFIED Python reference to 10^5 subframes with the Philox shim, seed 0
and 1. Stores one row (time, state, total_ues, reward, draws) per decision + the final PerfMonitor counters:
tests/golden/long_cfg1.npz  (build container only)."""
import json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from tools.refharness import RefCell
from tools.gen_golden import CFG1, DEFAULT_ACTION
T = 100000
res, counters = {}, []
for si, seed in enumerate([0, 1]):
    cell = RefCell(CFG1, seed)
    rec = cell.reset()
    rows = [(rec['time'], rec['state'], rec['total_ues'], rec['reward'], cell.rng.ctr)]
    while rec['time'] < T:
        rec = cell.step(DEFAULT_ACTION)
        rows.append((rec['time'], rec['state'], rec['total_ues'], rec['reward'], cell.rng.ctr))
    res[f'rows_{si}'] = np.asarray(rows, dtype=np.int64)
    counters.append(cell.counters())
    print(seed, len(rows), counters[-1])
meta = {'conf': CFG1, 'seeds': [0, 1], 'T': T, 'action': DEFAULT_ACTION, 'counters': counters}
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'long_cfg1.npz'), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **res)
